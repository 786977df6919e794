"""Recipe for ``oracle/_ref``: the reference's OWN CPU implementation of the flagging path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

The reference's CPU path is pure Python (``src/katsdpsigproc/rfi/host.py`` and the constant in
``rfi/__init__.py``; numpy + pandas only); its 2-D flagger ``rfi/twodflag.py`` is Python compiled
by numba (present in this image).  Where the reference checkout is present (the build
container: ``/root/reference``), this script places those two files, unmodified, under
``oracle/_ref/katsdpsigproc/rfi/`` -- git-ignored build output that travels to the GPU box
with the snapshot, like the compiled libraries -- so that ``bench.py``'s CPU legs can time the
unmodified ``FlaggerHost`` (``cpu_baseline.kind == "reference"``) and the parity block can
compare against it.  Nothing is copied into the tracked tree.  Without the checkout the
existing ``oracle/_ref`` (if any) is left alone and ``oracle.reference_host()`` falls back to
``None`` (callers then use the ``host_numpy`` port, kind "port").

    python oracle/make_ref.py [reference_root]
"""

from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref", "katsdpsigproc")
FILES = ("rfi/__init__.py", "rfi/host.py", "rfi/twodflag.py")


def make(reference_root: str = "/root/reference") -> bool:
    src = os.path.join(reference_root, "src", "katsdpsigproc")
    if not all(os.path.isfile(os.path.join(src, f)) for f in FILES):
        return False
    os.makedirs(os.path.join(DEST, "rfi"), exist_ok=True)
    # the package's own __init__ only resolves a version string (katversion); not needed
    with open(os.path.join(DEST, "__init__.py"), "w") as f:
        f.write("# placeholder package for the reference's rfi.host module (see oracle/make_ref.py)\n")
    for name in FILES:
        shutil.copyfile(os.path.join(src, name), os.path.join(DEST, name))
    return True


if __name__ == "__main__":
    ok = make(*sys.argv[1:2])
    print("oracle/_ref: " + ("reference host classes in place" if ok else "reference checkout not found"))
