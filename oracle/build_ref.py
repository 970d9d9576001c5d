"""Stage the reference's OWN implementation of the path under oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

The reference is pure Python, so "building" it means byte-compiling the two modules the hot path lives in —
`models/unet.py` (the U-Net, models/unet.py:40-92) and `metrics.py` (metrics.py:6-71) — from where they lie under
/root/reference into SOURCELESS `.pyc` files.  No reference source is copied into the repository: `oracle/_ref/` is
git-ignored (it still travels to the GPU box with the gpurun snapshot, like the built libclk.so), and only compiled
artefacts land there.  Used by `bench.py --impl reference` / `cpu_baseline` (kind "reference") to time the
reference's own module on the box's host cores, and by tests/test_oracle_pinned.py to pin the restatement.

    python -m oracle.build_ref          # /root/reference -> oracle/_ref/{models/__init__.pyc,models/unet.pyc,metrics.pyc}
"""
import importlib
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REFERENCE = "/root/reference"
FILES = ("models/__init__.py", "models/unet.py", "metrics.py")


def build_ref(reference=REFERENCE, out=OUT):
    """byte-compile the reference modules of the path; returns the output directory (None if there is no reference
    tree here, e.g. on the GPU box, where the prebuilt files are used)."""
    if not os.path.isdir(reference):
        return out if available(out) else None
    for rel in FILES:
        src = os.path.join(reference, rel)
        dst = os.path.join(out, rel + "c")          # legacy sourceless layout: pkg/mod.pyc next to where mod.py would be
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            py_compile.compile(src, cfile=dst, dfile=f"<reference>/{rel}", doraise=True)
    return out


def available(out=OUT):
    return all(os.path.exists(os.path.join(out, rel + "c")) for rel in FILES)


def load(out=OUT):
    """import the staged reference modules: returns (UNet class, metrics module).  Raises if they were not built."""
    if not available(out):
        raise RuntimeError("oracle/_ref is not built: run `python -m oracle.build_ref` where /root/reference exists")
    saved = {k: sys.modules.pop(k) for k in ("models", "models.unet", "metrics") if k in sys.modules}
    sys.path.insert(0, out)
    try:
        importlib.invalidate_caches()
        unet = importlib.import_module("models.unet")
        metrics = importlib.import_module("metrics")
    finally:
        sys.path.remove(out)
        for k in ("models", "models.unet", "metrics"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)
    return unet.UNet, metrics


if __name__ == "__main__":
    print(build_ref())
