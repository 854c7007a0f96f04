"""Golden fixture for a Chebyshev order above 7 (round 2: the transform takes 16 terms per launch, K > 7 is summed chunk by
chunk; the reference's loop, MagNetConv.py:213-240, has no limit).  Same method as make_golden.py.

    python tests/golden/make_golden_k9.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from make_golden import REF, magnet_case, nasty_graph  # noqa: E402

if __name__ == "__main__":
    ei, ew = nasty_graph(120, 700, seed=41)
    magnet_case("magnet_k9", REF["MagNetConv"], 120, 6, 4, 9, 0.15, 'sym', ei, ew, seed=42)
