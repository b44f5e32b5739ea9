"""Recipe for oracle/_ref: the UNMODIFIED reference, staged so that it can run on the GPU box's host cores.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference is thirteen plain Python files without a build system or
package metadata (no setup.py / pyproject.toml: ``pip install /root/reference`` has nothing to install), so "building"
it means packing its ``*.py`` byte for byte from the read-only mount into ONE archive, ``oracle/_ref/reference_py.tar.gz``
-- a git-ignored output directory (the sources never enter this repository's tree or history) that is NOT
gpurun-ignored, so the archive travels to the GPU box like the repo's own built ``.so``.  ``__graft_entry__.build()``
runs this whenever ``/root/reference`` exists; on the GPU box the archive is unpacked into a temporary directory outside
the repository when it is first needed.

``reference_modules()`` imports the staged files (or ``/root/reference`` itself) over ``oracle/pyg_shim.py`` -- stand-ins
for torch_geometric / xarray, which the image lacks -- and returns them by name.  Consumers: ``oracle/make_golden.py``
(fixtures), ``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` leg (``kind: "reference"``), and tests that
drive the reference's own training functions over the drop-in modules.  Never the product path.
"""
from __future__ import annotations

import contextlib
import glob
import hashlib
import importlib
import io
import os
import sys
import tarfile
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(REF_DST, "reference_py.tar.gz")
MODULES = ("graphBuilder", "model", "hybrid_model", "dataset", "embed_utils", "adaptive_scheduler", "featurePreprocessor",
           "train_hybrid_maml_v5")


def build(src=REF_SRC, archive=ARCHIVE):
    """Pack ``src/*.py`` unchanged into ``archive``.  Returns the member names (empty if ``src`` is absent)."""
    if not os.path.isdir(src):
        return []
    files = sorted(glob.glob(os.path.join(src, "*.py")))
    os.makedirs(os.path.dirname(archive), exist_ok=True)
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w:gz", format=tarfile.PAX_FORMAT) as tar:
        for f in files:
            info = tar.gettarinfo(f, arcname=os.path.basename(f))
            info.mtime, info.uid, info.gid, info.uname, info.gname = 0, 0, 0, "", ""  # reproducible archive
            with open(f, "rb") as fh:
                tar.addfile(info, fh)
    data = buf.getvalue()
    if not (os.path.exists(archive) and _members_equal(archive, files)):
        with open(archive, "wb") as fh:
            fh.write(data)
    return [os.path.basename(f) for f in files]


def _members_equal(archive, files):
    try:
        with tarfile.open(archive, "r:gz") as tar:
            names = {m.name: tar.extractfile(m).read() for m in tar.getmembers()}
    except (OSError, tarfile.TarError):
        return False
    return set(names) == {os.path.basename(f) for f in files} and all(
        open(f, "rb").read() == names[os.path.basename(f)] for f in files)


def _has_modules(d):
    return all(os.path.exists(os.path.join(d, m + ".py")) for m in MODULES)


_UNPACKED = None


def reference_dir():
    """Directory holding the unmodified reference files: the read-only mount where it exists (build container), else
    the staged archive unpacked once per process into a temporary directory; None if neither is there."""
    global _UNPACKED
    if _has_modules(REF_SRC):
        return REF_SRC
    if _UNPACKED is not None:
        return _UNPACKED
    if not os.path.exists(ARCHIVE):
        return None
    tag = hashlib.sha1(open(ARCHIVE, "rb").read()).hexdigest()[:12]
    d = os.path.join(tempfile.gettempdir(), f"wf_reference_{tag}")
    if not _has_modules(d):
        os.makedirs(d, exist_ok=True)
        with tarfile.open(ARCHIVE, "r:gz") as tar:
            for m in tar.getmembers():
                if m.isfile() and os.path.basename(m.name) == m.name and m.name.endswith(".py"):
                    with open(os.path.join(d, m.name), "wb") as fh:
                        fh.write(tar.extractfile(m).read())
    _UNPACKED = d if _has_modules(d) else None
    return _UNPACKED


def available():
    return reference_dir() is not None


def reference_modules(names=MODULES, quiet=True):
    """{name: module} of the unmodified reference files, imported over the third-party stand-ins.

    The reference's modules have top-level names (``model``, ``hybrid_model``, ...) that collide with nothing in this
    repository's package (its drop-ins live under ``weatherforecast_stgcn_maml_b200.``) unless a test has installed the
    drop-ins under those names; any such entries are removed first and restored by the caller if needed."""
    from oracle import pyg_shim

    d = reference_dir()
    if d is None:
        raise RuntimeError("the reference is not staged: run `python -m oracle.build_ref` where /root/reference exists")
    pyg_shim.install()
    for n in MODULES:
        m = sys.modules.get(n)
        if m is not None and os.path.dirname(os.path.abspath(getattr(m, "__file__", "") or "")) != d:
            del sys.modules[n]
    if d not in sys.path:
        sys.path.insert(0, d)
    out = {}
    sink = io.StringIO() if quiet else sys.stdout  # train_hybrid_maml_v5 prints a banner at import
    with contextlib.redirect_stdout(sink):
        for n in names:
            out[n] = importlib.import_module(n)
    return out


if __name__ == "__main__":
    files = build()
    print(f"[oracle/_ref] packed {len(files)} reference files into {ARCHIVE}" if files else "[oracle/_ref] /root/reference not found")
