"""Host-side mirror of the reference's conv-layer API (torch_geometric_signed_directed.nn):
same class names, constructor and forward signatures, parameter names and error behaviour,
backed by the sm_100a kernels in libpgsd_b200.so."""
from .magnet_conv import MagNetConv, MSConv
from .digcn_conv import DiGCNConv, DiGCN_InceptionBlock
from .sgcn_conv import SGCNConv
from .snea_conv import SNEAConv
from .mixed_path import Conv_Base, DIMPA
from .complex_relu import complex_relu_layer
from .dgcn_simpa import DGCNConv, SIMPA
from .sdr_layer import GATConv, SDRLayer
from .magnet_model import MagNet_node_classification
from .models import DiGCN_Inception_Block_node_classification, SGCN
from .signed_models import SDGNN, SiGAT

__all__ = ["MagNetConv", "MSConv", "DiGCNConv", "DiGCN_InceptionBlock", "SGCNConv", "SNEAConv",
           "Conv_Base", "DIMPA", "complex_relu_layer", "DGCNConv", "SIMPA", "GATConv", "SDRLayer",
           "MagNet_node_classification", "DiGCN_Inception_Block_node_classification", "SGCN",
           "SDGNN", "SiGAT"]
