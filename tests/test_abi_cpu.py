"""CPU: the C-ABI shared library loads without a GPU and exports every symbol include/mmb200.h declares."""
import ctypes
import os
import re

from multimodal_b200 import _abi, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mmb200.h")).read()
    return sorted(set(re.findall(r"^int (mmb_\w+)\(", src, flags=re.M)))


def test_library_loads_and_exports_header_symbols():
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mmb200.h but not exported"
    assert lib.mmb_version() >= 100


def test_ctypes_prototypes_cover_header():
    assert sorted(_abi.PROTOTYPES) == _declared()


def test_wgrad_split_heuristic():
    from multimodal_b200.ops import wgrad_splits

    for rows, cols, k in [(768, 3072, 201728), (3072, 768, 201728), (2304, 768, 201728), (768, 768, 201728), (512, 2048, 78848)]:
        s = wgrad_splits(rows, cols, k)
        tiles = -(-rows // 128) * -(-cols // 256) * s
        assert tiles >= 100 and s >= 1
