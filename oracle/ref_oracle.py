"""TEST INFRASTRUCTURE ONLY — opens oracle/_ref/libapd_ref.so (the reference's own APD.cu, compiled
unmodified for sm_100a by oracle/Makefile) through the same `Engine` class as the product.
Import from tests/, bench.py (--impl reference) and __graft_entry__.smoke() only."""
import os

from dvp_mvs_b200._lib import Engine

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_LIB = os.path.join(HERE, "_ref", "libapd_ref.so")


def available() -> bool:
    return os.path.exists(REFERENCE_LIB) and os.path.exists(os.path.join(HERE, "_ref", "libapd_ref_k2.so"))


def engine(width, height, num_src, params, device=0) -> Engine:
    return Engine(width, height, num_src, params, device=device, lib_path=REFERENCE_LIB, prefix="ref_")
