"""ORACLE (test infrastructure): compile the reference's own soft-NMS extension
(``/root/reference/lib/external/nms.pyx``, Cython) from the sources WHERE THEY LIE into
``oracle/_ref/`` — nothing is copied into the repository and the reference's build system
(``lib/external/setup.py`` / ``Makefile``) is not run.

    python -m oracle.build_ref        # -> oracle/_ref/nms*.so (+ the generated nms.c, both git-ignored)

Recipe: ``cython -3`` translates the .pyx to C in ``oracle/_ref/``; gcc builds the shared object against
this interpreter's headers and numpy.  One compatibility edit is applied to a TEMPORARY copy (never stored in
the repository): ``np.int_t`` (``nms.pyx:32``, inside the hard-``nms`` function that is not on this path) no
longer exists in numpy >= 2's Cython declarations and is spelled ``np.intp_t``; ``soft_nms_39`` (``:172-275``),
the function used as the oracle, is compiled exactly as written.  The build container only — ``/root/reference`` does not exist on
the GPU box; the compiled ``.so`` travels with the snapshot, and the golden vectors it produced
(``tests/golden/soft_nms.npz``, ``oracle/make_golden.py::gen_soft_nms``) are committed.
The rest of the reference's native code (DCNv2 against the removed THC API) is not buildable here
(DESIGN.md §3).
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

REF = "/root/reference/lib/external/nms.pyx"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def build(force: bool = False):
    """Returns the path of the built extension, or None when the reference tree is absent."""
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, "nms" + sysconfig.get_config_var("EXT_SUFFIX"))
    if os.path.exists(so) and not force:
        return so
    if not os.path.exists(REF):
        return None
    import numpy
    import tempfile
    c_file = os.path.join(OUT, "nms.c")
    with tempfile.TemporaryDirectory() as tmp:
        src = open(REF).read().replace("np.int_t", "np.intp_t")
        pyx = os.path.join(tmp, "nms.pyx")
        with open(pyx, "w") as f:
            f.write(src)
        subprocess.check_call([sys.executable, "-m", "cython", "-3", pyx, "-o", c_file])
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-Wno-cpp", "-Wno-unused-function",
                           "-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(),
                           c_file, "-o", so])
    return so


def load():
    """Import the compiled reference module (``nms.soft_nms_39`` ...), or None if unavailable."""
    so = build()
    if so is None:
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("nms", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force=True))
