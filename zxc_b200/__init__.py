"""zxc_b200 -- B200-native block codec behind the ZXC C API.

The product is the C-ABI shared library ``zxc_b200/lib/libzxc.so.4`` (built by
``__graft_entry__.build()`` from ``zxc_b200/csrc``); this package is only a thin
ctypes loader so Python callers (tests, bench, torch.distributed drivers) reach the
same entry points a C caller links against.  Importing it fails loudly when the
library has not been built -- there is no Python or CPU fallback codec.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libzxc.so.4")

if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` first")

lib = ctypes.CDLL(LIB_PATH)
lib.zxc_version_string.restype = ctypes.c_char_p
__version__ = lib.zxc_version_string().decode()


def device_count():
    return int(lib.zxc_b200_device_count())
