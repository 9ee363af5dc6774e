"""Build recipe for the oracle (TEST INFRASTRUCTURE, see oracle/__init__.py).

Two artefacts, both git-ignored, both travelling to the GPU box with gpurun:

* ``oracle/liboracle_port.so``  -- gcc build of ``oracle/pypore_oracle.c`` (the
  CPU restatement of the hot path).
* ``oracle/_ref/cparsers*.so``  -- the reference's own ``PyPore/cparsers.pyx``,
  cythonised and compiled UNMODIFIED from where it lies under /root/reference
  (SURVEY.md App. B).  Outputs only go into ``oracle/_ref``; no reference source
  is copied into the repository (the generated ``cparsers.c`` lives in
  ``oracle/_ref`` too, which is git-ignored).  Skipped when /root/reference is
  absent (the GPU box): the prebuilt file is used there.

Flags: plain ``-O2``, ``-ffp-contract=off``; never ``-march=native``/``-mfma``
(FMA contraction changes the reference's rounding).
"""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE = os.environ.get("PYPORE_REFERENCE", "/root/reference")
PORT_SO = os.path.join(HERE, "liboracle_port.so")


def _newer(target, *sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def build_port(force=False):
    src = os.path.join(HERE, "pypore_oracle.c")
    if not force and _newer(PORT_SO, src):
        return PORT_SO
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC",
           src, "-o", PORT_SO, "-lm"]
    subprocess.check_call(cmd)
    return PORT_SO


def ref_so_path():
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return os.path.join(REF_DIR, "PyPore", "cparsers" + suffix)


def build_ref(force=False):
    """Compile the unmodified reference cparsers.pyx into oracle/_ref/PyPore/."""
    pyx = os.path.join(REFERENCE, "PyPore", "cparsers.pyx")
    out = ref_so_path()
    if not os.path.exists(pyx):
        return out if os.path.exists(out) else None
    if not force and _newer(out, pyx):
        return out
    import numpy
    pkg = os.path.join(REF_DIR, "PyPore")
    os.makedirs(pkg, exist_ok=True)
    # cython reads the .pyx in place and writes the generated C into oracle/_ref.
    c_file = os.path.join(pkg, "cparsers.c")
    subprocess.check_call([sys.executable, "-m", "cython", "-2", pyx, "-o", c_file])
    inc = sysconfig.get_paths()["include"]
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-w",
           "-I" + inc, "-I" + numpy.get_include(), c_file, "-o", out, "-lm"]
    subprocess.check_call(cmd)
    return out


def build_all(force=False):
    port = build_port(force)
    ref = build_ref(force)
    return port, ref


if __name__ == "__main__":
    if "--clean" in sys.argv:
        shutil.rmtree(REF_DIR, ignore_errors=True)
        if os.path.exists(PORT_SO):
            os.remove(PORT_SO)
    print(build_all(force="--force" in sys.argv))
