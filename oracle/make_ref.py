"""Recipe for oracle/_ref/: a verbatim, git-ignored copy of the reference's own package.  TEST INFRASTRUCTURE ONLY.

The reference (catniplab/vjf) is pure Python + torch, so "building" it means copying its ten source files
from where they lie (/root/reference/vjf, read-only, present in the build container only) into
oracle/_ref/vjf/.  oracle/_ref/ is listed in .gitignore (no reference source enters the history) but NOT in
.gpurunignore, so it travels to the GPU box with the snapshot like the built .so files.  It is used by

  * bench.py --impl reference and bench.py's cpu_baseline leg (kind "reference"): the UNMODIFIED
    vjf.model.VJF.filter timed on the host cores, and
  * tests/golden/make_golden.py (which can also import /root/reference directly).

Nothing under vjf_b200/ imports it.  Run:  python oracle/make_ref.py   (idempotent; __graft_entry__.build() calls it
when /root/reference is present).
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("VJF_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(DST, "vjf", "model.py"))


def make(verbose=False):
    src = os.path.join(REF_SRC, "vjf")
    if not os.path.isdir(src):
        if verbose:
            print(f"{src} is absent (GPU box?): keeping whatever is in {DST}")
        return available()
    os.makedirs(os.path.join(DST, "vjf"), exist_ok=True)
    for f in sorted(os.listdir(src)):
        if not f.endswith(".py"):
            continue
        a, b = os.path.join(src, f), os.path.join(DST, "vjf", f)
        if not (os.path.exists(b) and filecmp.cmp(a, b, shallow=False)):
            shutil.copyfile(a, b)
            if verbose:
                print("copied", f)
    return available()


def import_reference():
    """Import the unmodified reference package from oracle/_ref (or /root/reference when present)."""
    for root in (DST, REF_SRC):
        if os.path.isfile(os.path.join(root, "vjf", "model.py")):
            if root not in sys.path:
                sys.path.insert(0, root)
            import vjf.model as ref_model  # noqa: F401
            return ref_model
    raise ImportError("the reference package is neither in oracle/_ref nor in /root/reference: run python oracle/make_ref.py in the build container")


if __name__ == "__main__":
    ok = make(verbose=True)
    print("oracle/_ref", "ready" if ok else "MISSING")
    sys.exit(0 if ok else 1)
