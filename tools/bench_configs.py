"""Timings of the other BASELINE configs (parity-test cases, not bench lines): C3 DiGCN inception
block 500k nodes / 2x10M edges / 128 bf16, C4 SGCN 2M nodes / 40M signed entries / 64 fp32, and
DIMPA.  Writes gpurun_out/configs.jsonl with algorithmic GB/s per layer (SURVEY §8d formula)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_geometric_signed_directed_b200 import nn, ops, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
os.makedirs("gpurun_out", exist_ok=True)
fh = open("gpurun_out/configs.jsonl", "a")
PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0


def emit(**kw):
    print(json.dumps(kw), flush=True)
    fh.write(json.dumps(kw) + "\n")
    fh.flush()


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def breakdown(fn, iters=3):
    """Per-kernel-family device time of one call (CUDA events around every launch, ops.TIMING)."""
    ops.TIMING = []
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    t, ops.TIMING = ops.TIMING, None
    out = {}
    for name, a, b in t:
        out[name] = out.get(name, 0.0) + a.elapsed_time(b) / iters
    return {k: round(v, 4) for k, v in out.items()}


with torch.no_grad():
    # ---- C3: DiGCN_InceptionBlock, 500k nodes, two 10M-nnz operators, 128 features, bf16
    n, e, f = 500_000, 10_000_000, 128
    ei1, _ = synthetic.dsbm_edges(n, 3, num_edges=e, seed=1, device=dev)
    ei2, _ = synthetic.dsbm_edges(n, 3, num_edges=e, seed=2, device=dev)
    w1, w2 = synthetic.sym_norm_weights(ei1, n), synthetic.sym_norm_weights(ei2, n)
    x = (torch.rand(n, f, device=dev) * 2 - 1).bfloat16()
    blk = nn.DiGCN_InceptionBlock(f, f).to(dev)
    blk(x, ei1, w1, ei2, w2)
    nnz = ei1.size(1) + ei2.size(1)
    b_alg = nnz * 8 + 2 * (n + 1) * 4 + nnz * f * 2 + n * f * 2 * (1 + 3 + 2 + 2)   # x, buf(3), 2 gathers src, x1, x2
    ms = timeit(lambda: blk(x, ei1, w1, ei2, w2))
    emit(config="C3 DiGCN_InceptionBlock 500k/2x10M/128 bf16", ms=ms, edges_per_s=nnz / ms * 1e3,
         alg_gb=b_alg / 1e9, alg_gbs=b_alg / ms / 1e6, frac=b_alg / ms / 1e6 / PEAK,
         kernels_ms=breakdown(lambda: blk(x, ei1, w1, ei2, w2)))
    # C3 as a model: DiGCN_Inception_Block_node_classification (3 blocks, x0+x1+x2 fused into the aggregation epilogues)
    mdl = nn.DiGCN_Inception_Block_node_classification(f, f, 16, dropout=0.5).to(dev).to(torch.bfloat16).eval()
    mdl(x, (ei1, ei2), (w1, w2))
    ms = timeit(lambda: mdl(x, (ei1, ei2), (w1, w2)))
    emit(config="C3 DiGCN_Inception_Block_node_classification 3 blocks 500k/2x10M/128->128->16 bf16 (inference)", ms=ms,
         edges_per_s=3 * nnz / ms * 1e3, kernels_ms=breakdown(lambda: mdl(x, (ei1, ei2), (w1, w2))))
    del ei1, ei2, w1, w2, x, blk, mdl
    torch.cuda.empty_cache()

    # ---- C4: SGCN (in=64, out=64 -> conv1 64->32|32, conv2 (32|32)->32|32), 2M nodes, 40M signed entries
    n = 2_000_000
    pos, neg, _ = synthetic.ssbm_edges(n, 3, num_entries=40_000_000, eta=0.1, seed=0, device=dev)
    x = torch.randn(n, 64, device=dev)
    c1 = nn.SGCNConv(64, 32, first_aggr=True).to(dev)
    c2 = nn.SGCNConv(32, 32, first_aggr=False).to(dev)
    z1 = torch.tanh(c1(x, pos, neg))
    c2(z1, pos, neg)
    nnz = pos.size(1) + neg.size(1)
    # col index per entry, two row_ptr arrays, one 64-wide gather per entry, x read once, out written once
    # (intermediates the implementation may write are NOT algorithmic bytes)
    b1 = nnz * 4 + 2 * (n + 1) * 4 + nnz * 64 * 4 + n * 64 * 4 * 2
    ms1 = timeit(lambda: c1(x, pos, neg))
    ms2 = timeit(lambda: c2(z1, pos, neg))
    emit(config="C4 SGCNConv layer 1 2M/40M/64 fp32", ms=ms1, edges_per_s=nnz / ms1 * 1e3, alg_gb=b1 / 1e9,
         alg_gbs=b1 / ms1 / 1e6, frac=b1 / ms1 / 1e6 / PEAK, kernels_ms=breakdown(lambda: c1(x, pos, neg)))
    emit(config="C4 SGCNConv layer 2 2M/40M/(32|32) fp32", ms=ms2, edges_per_s=nnz / ms2 * 1e3, alg_gb=b1 / 1e9,
         alg_gbs=b1 / ms2 / 1e6, frac=b1 / ms2 / 1e6 / PEAK, kernels_ms=breakdown(lambda: c2(z1, pos, neg)))
    # SNEAConv (rows a8): same SSBM graph, 64 -> 32|32 first layer and (32|32) -> 32|32 deep layer
    s1 = nn.SNEAConv(64, 32, first_aggr=True).to(dev)
    s2 = nn.SNEAConv(32, 32, first_aggr=False).to(dev)
    zs = s1(x, pos, neg)
    s2(zs, pos, neg)
    # Algorithmic bytes of the attention kernels: edge softmax = per stored entry the column index and
    # ONE scalar score gather (4 + 4 B; the 32-byte sector a scalar gather really moves is traffic, not algorithm),
    # per row and type row_ptr + s_dst (8 B), plus what the mode writes: SNEAConv (mode A) reads the target's feature
    # row per type and writes the output row; GAT-style (mode B) writes alpha per entry (4 B).
    def softmax_bytes_snea(plans, feat):
        return sum(p.nnz * 8 + p.n_dst * (8 + feat * 4) for p in plans) + plans[0].n_dst * feat * 4

    p_pos, p_neg = s1._plan_for(pos, n, True), s1._plan_for(neg, n, True)
    p_neg2 = s2._plan_for(neg, n, False)
    att_bytes = {"SNEAConv layer 1 2M/40M/64 fp32": softmax_bytes_snea([p_pos], 32) + softmax_bytes_snea([p_neg], 32),
                 "SNEAConv layer 2 2M/40M/(32|32) fp32": 2 * softmax_bytes_snea([p_pos, p_neg2], 32)}
    for nm, fn in (("SNEAConv layer 1 2M/40M/64 fp32", lambda: s1(x, pos, neg)),
                   ("SNEAConv layer 2 2M/40M/(32|32) fp32", lambda: s2(zs, pos, neg))):
        ms_ = timeit(fn)
        kb = breakdown(fn)
        sm = kb.get("edge_softmax", 0.0)
        emit(config=nm, ms=ms_, edges_per_s=nnz / ms_ * 1e3, kernels_ms=kb,
             edge_softmax={"alg_gb": att_bytes[nm] / 1e9, "ms": sm,
                           "frac": att_bytes[nm] / sm / 1e6 / PEAK if sm else None})
    del s1, s2, zs
    # SDRLayer (row a9): 4 GATConv over the pos/neg in/out lists of the same graph + MLP, 64 -> 64
    sdr = nn.SDRLayer(64, 64, [pos, pos.flip(0), neg, neg.flip(0)]).to(dev)
    sdr(x)
    ms_ = timeit(lambda: sdr(x))
    kb = breakdown(lambda: sdr(x))
    gp = [a._plan_for(e_, n) for a, e_ in zip(sdr.aggs, sdr.edge_lists)]
    b_sm = sum(p.nnz * (8 + 4) + p.n_dst * 8 for p in gp)                       # mode B: + alpha written per entry
    b_ag = sum(p.nnz * (4 + 4 + 64 * 4) + (p.n_dst + 1) * 4 + 2 * p.n_dst * 64 * 4 for p in gp)   # val = alpha, z and y
    emit(config="SDRLayer 4 x GATConv 2M/2x40M/64 fp32", ms=ms_, edges_per_s=2 * nnz / ms_ * 1e3, kernels_ms=kb,
         edge_softmax={"alg_gb": b_sm / 1e9, "ms": kb.get("edge_softmax"),
                       "frac": b_sm / kb["edge_softmax"] / 1e6 / PEAK if kb.get("edge_softmax") else None},
         gat_aggregation={"alg_gb": b_ag / 1e9, "ms": kb.get("spmm"),
                          "frac": b_ag / kb["spmm"] / 1e6 / PEAK if kb.get("spmm") else None})
    del sdr
    torch.cuda.empty_cache()
    # C4 as a model: SGCN(in=64, out=64, 2 layers), tanh fused into each layer's transform
    es = torch.cat([torch.cat([pos.t(), torch.ones(pos.size(1), 1, dtype=torch.long, device=dev)], 1),
                    torch.cat([neg.t(), -torch.ones(neg.size(1), 1, dtype=torch.long, device=dev)], 1)], 0)
    sg = nn.SGCN(n, es, in_dim=64, out_dim=64, layer_num=2, init_emb=x).to(dev).eval()
    del es
    sg()
    ms = timeit(lambda: sg())
    emit(config="C4 SGCN model 2 layers 2M/40M/64 fp32 (inference)", ms=ms, edges_per_s=2 * nnz / ms * 1e3,
         alg_gb=2 * b1 / 1e9, alg_gbs=2 * b1 / ms / 1e6, frac=2 * b1 / ms / 1e6 / PEAK, kernels_ms=breakdown(lambda: sg()))
    del pos, neg, x, z1, sg
    torch.cuda.empty_cache()

    # ---- C2 as a whole model: MagNet_node_classification, 2 layers, 1M nodes / 20M edges / 64 hidden
    n, f, lab = 1_000_000, 64, 10
    ei, _ = synthetic.dsbm_edges(n, 3, num_edges=20_000_000, seed=0, device=dev)
    x = torch.rand(n, f, device=dev) * 2 - 1
    model = nn.MagNet_node_classification(f, hidden=f, q=0.25, K=1, label_dim=lab, activation=True, layer=2,
                                          cached=True).to(dev).eval()
    model(x, x, ei)
    p0 = model.Chebs[0]._plan
    nnz = int(p0.nnz)
    layer = nnz * (4 + 8 + 2 * f * 4) + (n + 1) * 4 + 4 * n * f * 4      # SURVEY 8d per layer, as bench.py
    layer_shared = nnz * (4 + 8 + f * 4) + (n + 1) * 4 + 4 * n * f * 4   # x_real is x_imag: one gather per entry
    tail = n * 2 * f * 4 + n * lab * 4 * 3
    # (a) the way the reference's example calls it (examples/magnet_node.py:61-62: X_real = X_img = data.x):
    #     layer 1 gathers each neighbour row once for both operators
    b = layer_shared + layer + tail
    ms = timeit(lambda: model(x, x, ei))
    emit(config="C2 MagNet_node_classification 2 layers 1M/20M/64 fp32 (inference, X_real is X_img)", ms=ms,
         edges_per_s=2 * ei.size(1) / ms * 1e3, alg_gb=b / 1e9, alg_gbs=b / ms / 1e6, frac=b / ms / 1e6 / PEAK,
         fused_layer=ops.FUSED_LAYER, kernels_ms=breakdown(lambda: model(x, x, ei)))
    # (b) two distinct input tensors: the general path in both layers
    x2 = x.clone()
    b = 2 * layer + tail
    ms = timeit(lambda: model(x, x2, ei))
    emit(config="C2 MagNet_node_classification 2 layers 1M/20M/64 fp32 (inference, distinct X_real / X_img)", ms=ms,
         edges_per_s=2 * ei.size(1) / ms * 1e3, alg_gb=b / 1e9, alg_gbs=b / ms / 1e6, frac=b / ms / 1e6 / PEAK,
         fused_layer=ops.FUSED_LAYER, kernels_ms=breakdown(lambda: model(x, x2, ei)))
    del model, x, x2, ei, p0
    torch.cuda.empty_cache()

    # ---- DIMPA hop=2 on 1M nodes / 20M edges / 64
    n = 1_000_000
    ei, _ = synthetic.dsbm_edges(n, 3, num_edges=20_000_000, seed=0, device=dev)
    ew = torch.rand(ei.size(1), device=dev) + 0.5
    xs, xt = torch.rand(n, 64, device=dev), torch.rand(n, 64, device=dev)
    dm = nn.DIMPA(hop=2).to(dev)
    dm(xs, xt, ei, ew)
    ms = timeit(lambda: dm(xs, xt, ei, ew))
    b = 4 * (ei.size(1) * (8 + 256) + (n + 1) * 4 + n * 256 * 2) + 4 * n * 256 * 3
    emit(config="DIMPA hop=2 1M/20M/64 fp32", ms=ms, edges_per_s=4 * ei.size(1) / ms * 1e3, alg_gb=b / 1e9,
         alg_gbs=b / ms / 1e6, frac=b / ms / 1e6 / PEAK, kernels_ms=breakdown(lambda: dm(xs, xt, ei, ew)))

# ---- C2 training step (forward + nll loss + backward through the same kernels), same model
n, f, lab = 1_000_000, 64, 10
ei, _ = synthetic.dsbm_edges(n, 3, num_edges=20_000_000, seed=0, device=dev)
x = torch.rand(n, f, device=dev) * 2 - 1
yl = torch.randint(0, lab, (n,), device=dev)
model = nn.MagNet_node_classification(f, hidden=f, q=0.25, K=1, label_dim=lab, activation=True, layer=2,
                                      cached=True).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3)


def train_step():
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.nll_loss(model(x, x, ei), yl)
    loss.backward()
    opt.step()


ms = timeit(train_step)
emit(config="C2 MagNet_node_classification 2 layers 1M/20M/64 fp32 (training step, Adam)", ms=ms,
     edges_per_s=2 * ei.size(1) / ms * 1e3)
