import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")


def _ensure_built():
    from approximated_personalized_pagerank_b200 import _lib
    import oracle_bindings as ob
    if not _lib.LIB_PATH.exists() or not ob.ORACLE_PATH.exists():
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session", autouse=True)
def built():
    _ensure_built()


@pytest.fixture(scope="session")
def ppr():
    import approximated_personalized_pagerank_b200 as m
    return m


@pytest.fixture(scope="session")
def ob():
    import oracle_bindings
    return oracle_bindings


def golden_cases(prefix):
    return sorted(p.stem for p in (ROOT / "tests" / "golden").glob(f"{prefix}_*.npz"))


def load_golden(name):
    import numpy as np
    from approximated_personalized_pagerank_b200.graphs import CSRGraph
    z = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
    g = CSRGraph(z["row_ptr"].astype(np.int64), z["col"].astype(np.int32), None)
    return g, z


requires_ref = pytest.mark.skipif(not (ROOT / "oracle" / "_ref" / "libppr_ref.so").exists(),
                                  reason="oracle/_ref/libppr_ref.so not built (needs /root/reference)")
