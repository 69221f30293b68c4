"""Stages the UNMODIFIED reference as byte-compiled, source-less modules under oracle/_ref/ so that it can be TIMED on the GPU
box's host cores (bench.py --impl reference, cpu_baseline.kind = "reference").  TEST / BENCH INFRASTRUCTURE ONLY.

/root/reference is a pure-Python script tree without setup.py (it cannot be pip-installed), and it does not exist on the GPU
box.  This recipe compiles every module WHERE IT LIES under /root/reference with the image's own interpreter (py_compile, the
Python analogue of `gcc file.c -o oracle/_ref/x.so`) and writes ONLY the resulting .pyc files into oracle/_ref/<package path>/
(git-ignored, travels to the GPU box with the snapshot like a built .so).  No reference source is copied into the repository.
The four import stand-ins the reference needs here (anytree, gym, colorama, multiprocessing_logging: not installed, no network)
are this repo's own oracle/_shims/.

    python -m oracle.build_ref          (also run by __graft_entry__.build() when /root/reference is present)
"""
import os
import py_compile
import shutil
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = "/root/reference"
REF_DIR = os.path.join(_HERE, "_ref")
SHIMS = os.path.join(_HERE, "_shims")
PACKAGES = ("games", "rl_utils")
SUFFIX = ".refpyc"   # not ".pyc": snapshot tools (the GPU runner among them) drop *.pyc / __pycache__ as build litter
STAMP = os.path.join(REF_DIR, "BUILT_FROM")


def available():
    """True when the staged reference can be imported (this container after build(); the GPU box via the snapshot)."""
    return os.path.exists(os.path.join(REF_DIR, "games", "algos", "mcts" + SUFFIX))


def build(force=False):
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "games")):
        return REF_DIR if available() else None          # GPU box: use what travelled
    tag = f"{SUFFIX} {REFERENCE_ROOT} python {sys.version_info[0]}.{sys.version_info[1]} magic {py_compile.importlib.util.MAGIC_NUMBER.hex()}"
    if not force and available() and os.path.exists(STAMP) and open(STAMP).read() == tag:
        return REF_DIR
    shutil.rmtree(REF_DIR, ignore_errors=True)
    n = 0
    for pkg in PACKAGES:
        for dirpath, dirnames, files in os.walk(os.path.join(REFERENCE_ROOT, pkg)):
            dirnames[:] = [d for d in dirnames if d != "__pycache__"]
            rel = os.path.relpath(dirpath, REFERENCE_ROOT)
            for fn in files:
                if fn.endswith(".py"):
                    out = os.path.join(REF_DIR, rel, fn[:-3] + SUFFIX)
                    os.makedirs(os.path.dirname(out), exist_ok=True)
                    # dfile: the path shown in tracebacks stays the reference's own
                    py_compile.compile(os.path.join(dirpath, fn), cfile=out, dfile=os.path.join(REFERENCE_ROOT, rel, fn), doraise=True,
                                       invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
                    n += 1
    with open(STAMP, "w") as f:
        f.write(tag)
    return REF_DIR


def import_paths():
    """sys.path entries (in order) that make `import games.algos.mcts` resolve to the staged reference.  Also registers the
    path hook that loads the staged byte-code files (a sourceless loader for SUFFIX, consulted for oracle/_ref only)."""
    import importlib.machinery as m
    if not getattr(import_paths, "_hooked", False):
        file_hook = m.FileFinder.path_hook((m.SourcelessFileLoader, [SUFFIX]))

        def ref_hook(path):
            if not os.path.abspath(path).startswith(REF_DIR):
                raise ImportError("not the staged reference")
            return file_hook(path)
        sys.path_hooks.insert(0, ref_hook)
        sys.path_importer_cache.clear()
        import_paths._hooked = True
    return [REF_DIR, SHIMS]


if __name__ == "__main__":
    print(build(force=True), "available:", available())
