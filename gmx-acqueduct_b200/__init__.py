"""gmx-acqueduct_b200: B200-native hydrogen-bond pair search + energy path of
gmx-waternetwork.  The product is libwn_b200.so (csrc/, C ABI in include/wn_b200.h)
and the C++ host mirror of the reference interface (include/*.hpp, host/).
This Python package only holds the ctypes binding used by tests and bench.py.

The directory name carries a hyphen (it follows the reference repository's name),
so load it with `wn_pkg.load_capi()` from the repository root or via importlib."""
