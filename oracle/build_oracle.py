"""TEST INFRASTRUCTURE ONLY -- builds the checkers:

  build_c()    gcc -> oracle/_build/libmas_oracle.so from oracle/mas_oracle.c (the CPU restatement);
  build_ref()  when /root/reference is present (the build container): cythonize the UNMODIFIED
               Matcha-TTS/matcha/utils/monotonic_align/core.pyx where it lies and compile it into oracle/_ref/
               (git-ignored, travels to the GPU box).  No reference source is copied into the repo: only the generated
               C file and the extension module land in oracle/_ref/.

    python -m oracle.build_oracle
"""
import glob
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PYX = "/root/reference/Matcha-TTS/matcha/utils/monotonic_align/core.pyx"
REF_DIR = os.path.join(HERE, "_ref")
BUILD_DIR = os.path.join(HERE, "_build")


def build_c() -> str:
    src, out = os.path.join(HERE, "mas_oracle.c"), os.path.join(BUILD_DIR, "libmas_oracle.so")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(BUILD_DIR, exist_ok=True)
        # -ffp-contract=off: one rounded float32 add per cell, exactly as the Cython code compiles
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", out], check=True)
    return out


def ref_module_path():
    hits = glob.glob(os.path.join(REF_DIR, "mas_ref_core*.so"))
    return hits[0] if hits else None


def build_ref():
    """-> path of the compiled reference extension, or None when the reference tree / cython are unavailable."""
    if ref_module_path() and (not os.path.exists(REF_PYX) or os.path.getmtime(ref_module_path()) >= os.path.getmtime(REF_PYX)):
        return ref_module_path()
    if not os.path.exists(REF_PYX) or shutil.which("gcc") is None:
        return None
    try:
        import Cython  # noqa: F401
        import numpy as np
    except ImportError:
        return None
    os.makedirs(REF_DIR, exist_ok=True)
    c_file = os.path.join(REF_DIR, "mas_ref_core.c")
    # the module is renamed (-o + --module-name) so that it cannot shadow anything; the source is read in place
    r = subprocess.run([sys.executable, "-m", "cython", "-3", "--module-name", "mas_ref_core", REF_PYX, "-o", c_file],
                       capture_output=True, text=True)
    if r.returncode != 0 or not os.path.exists(c_file):
        return None
    out = os.path.join(REF_DIR, "mas_ref_core" + sysconfig.get_config_var("EXT_SUFFIX"))
    cmd = ["gcc", "-O2", "-shared", "-fPIC", "-fopenmp", "-I" + sysconfig.get_paths()["include"], "-I" + np.get_include(), c_file, "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return out if r.returncode == 0 else None


def load_ref():
    """Import the compiled reference extension from oracle/_ref/ (None when it was never built)."""
    p = ref_module_path()
    if p is None:
        return None
    import importlib.util

    spec = importlib.util.spec_from_file_location("mas_ref_core", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print("oracle C restatement:", build_c())
    print("reference build     :", build_ref())
