"""Stage the reference's OWN implementation of the path under oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

The reference is pure Python, so "building" it means byte-compiling the two modules the hot path lives in —
`models/unet.py` (the U-Net, models/unet.py:40-92) and `metrics.py` (metrics.py:6-71) — from where they lie under
/root/reference into SOURCELESS bytecode files (`*.refbin` = the `.pyc` py_compile writes; a neutral suffix, because snapshot tools
tend to drop `*.pyc`).  No reference source is copied into the repository: `oracle/_ref/` is
git-ignored (it still travels to the GPU box with the gpurun snapshot, like the built libclk.so), and only compiled
artefacts land there.  Used by `bench.py --impl reference` / `cpu_baseline` (kind "reference") to time the
reference's own module on the box's host cores, and by tests/test_oracle_pinned.py to pin the restatement.

    python -m oracle.build_ref          # /root/reference -> oracle/_ref/{models.unet,metrics}.refbin
"""
import marshal
import os
import py_compile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REFERENCE = "/root/reference"
FILES = ("models/unet.py", "metrics.py")


def _dst(out, rel):
    return os.path.join(out, rel[:-3].replace("/", ".") + ".refbin")


def build_ref(reference=REFERENCE, out=OUT):
    """byte-compile the reference modules of the path; returns the output directory (None if there is no reference
    tree here, e.g. on the GPU box, where the prebuilt files are used)."""
    if not os.path.isdir(reference):
        return out if available(out) else None
    for rel in FILES:
        src = os.path.join(reference, rel)
        dst = _dst(out, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            py_compile.compile(src, cfile=dst, dfile=f"<reference>/{rel}", doraise=True)
    return out


def available(out=OUT):
    return all(os.path.exists(_dst(out, rel)) for rel in FILES)


def _exec(out, rel, name):
    with open(_dst(out, rel), "rb") as f:
        data = f.read()
    code = marshal.loads(data[16:])            # 16-byte .pyc header (magic, flags, mtime, size), then the code object
    mod = types.ModuleType(name)
    mod.__file__ = f"<reference>/{rel}"
    exec(code, mod.__dict__)
    return mod


def load(out=OUT):
    """execute the staged reference modules: returns (UNet class, metrics module).  Raises if they were not built
    (or were built by another Python version: the bytecode is only valid for the interpreter that compiled it —
    this image's on both sides)."""
    if not available(out):
        raise RuntimeError("oracle/_ref is not built: run `python -m oracle.build_ref` where /root/reference exists")
    unet = _exec(out, "models/unet.py", "_clk_reference_models_unet")
    metrics = _exec(out, "metrics.py", "_clk_reference_metrics")
    return unet.UNet, metrics


if __name__ == "__main__":
    print(build_ref())
