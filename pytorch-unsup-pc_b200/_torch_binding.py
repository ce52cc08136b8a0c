"""Build / load the optional C++ autograd binding (csrc/torch_binding.cpp -> lib/dpc_b200_torch.so).

The binding calls the same C ABI as the ctypes path; it exists so that the eager forward and
backward run without the Python interpreter.  ``DPC_B200_NO_TORCH_BINDING=1`` disables it."""
import importlib.util
import os
import shutil

_HERE = os.path.dirname(os.path.abspath(__file__))
NAME = "dpc_b200_torch"
SO_PATH = os.path.join(_HERE, "lib", NAME + ".so")
_mod = False     # False: not tried yet; None: unavailable


def build(verbose=False):
    """Compile the binding with torch's own extension builder (g++, no nvcc needed) against the
    already built lib/libdpc_b200.so and place it next to it."""
    import torch.utils.cpp_extension as ce
    build_dir = os.path.join(_HERE, "build", "torch_binding")
    os.makedirs(build_dir, exist_ok=True)
    lib_dir = os.path.join(_HERE, "lib")
    ce.load(name=NAME, sources=[os.path.join(_HERE, "csrc", "torch_binding.cpp")],
            extra_cflags=["-O2"], extra_ldflags=["-L" + lib_dir, "-ldpc_b200", "-Wl,-rpath,'$$ORIGIN'",
                           "-Wl,-rpath,'$$ORIGIN/../../lib'"],   # from lib/ and from the build dir
            build_directory=build_dir, with_cuda=True, is_python_module=False, verbose=verbose)
    shutil.copyfile(os.path.join(build_dir, NAME + ".so"), SO_PATH)
    return SO_PATH


def load():
    """The binding module, or None when it is not built / disabled / does not match the library."""
    global _mod
    if _mod is not False:
        return _mod
    _mod = None
    if os.environ.get("DPC_B200_NO_TORCH_BINDING") or not os.path.isfile(SO_PATH):
        return None
    from . import _lib
    lib = _lib.load()                      # the C ABI library first: the binding links against it
    if os.path.realpath(_lib.LIB_PATH) != os.path.realpath(os.path.join(_HERE, "lib", "libdpc_b200.so")):
        return None                        # an A/B variant library is in use: stay on ctypes
    import torch  # noqa: F401  (libtorch symbols)
    spec = importlib.util.spec_from_file_location(NAME, SO_PATH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if mod.abi_version() != lib.dpc_version():
        return None
    _mod = mod
    return mod
