"""`python setup.py build_ext --inplace` -- the build step the reference's README.md:1-2 names (there it compiles
TGL's sampler_core; the reference ships neither that setup.py nor the C++ source).  Here it builds,
in-tree:

  1. tgb-tgn-dgl_b200/lib/libtgn_b200.so   every CUDA kernel of the hot path for sm_100a behind the C-ABI of
                                            include/tgn_b200.h (make -C tgb-tgn-dgl_b200/csrc; nvcc cross-compiles
                                            without a GPU);
  2. tgb-tgn-dgl_b200/tgn_b200/_tgn_torch*.so   the torch C++ extension over that C-ABI (csrc_ext/tgn_torch.cpp,
                                            torch.ops.tgn.*), linked against (1) with an $ORIGIN-relative rpath.

Nothing is installed into site-packages: the package is used from the tree (and travels to the GPU box that way).
"""
import os
import subprocess
import sys

from setuptools import setup

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(HERE, "tgb-tgn-dgl_b200")


def build_cabi():
    env = dict(os.environ)
    env["PATH"] = "/usr/local/cuda/bin:" + env.get("PATH", "")
    subprocess.check_call(["make", "-C", os.path.join(PKG, "csrc"), "-j8"], env=env)


def torch_extension():
    from torch.utils.cpp_extension import CppExtension
    return CppExtension(
        name="_tgn_torch",
        sources=[os.path.join("tgb-tgn-dgl_b200", "csrc_ext", "tgn_torch.cpp")],
        include_dirs=["/usr/local/cuda/include", os.path.join(HERE, "include")],
        library_dirs=[os.path.join(PKG, "lib"), "/usr/local/cuda/lib64"],
        libraries=["tgn_b200", "c10_cuda"],
        extra_compile_args=["-O2", "-g0", "-std=c++17"],
        extra_link_args=["-Wl,-rpath,$ORIGIN/../lib"])


if __name__ == "__main__":
    from torch.utils.cpp_extension import BuildExtension

    class Build(BuildExtension):
        def run(self):
            build_cabi()
            super().run()

        def get_ext_fullpath(self, ext_name):     # in-tree target: tgb-tgn-dgl_b200/tgn_b200/
            name = os.path.basename(super().get_ext_fullpath(ext_name))
            return os.path.join(PKG, "tgn_b200", name)

        def copy_extensions_to_source(self):      # already linked into the tree by get_ext_fullpath
            pass

    setup(name="tgn_b200", version="0.2", ext_modules=[torch_extension()],
          cmdclass={"build_ext": Build.with_options(use_ninja=False)}, script_args=sys.argv[1:] or ["build_ext", "--inplace"])
