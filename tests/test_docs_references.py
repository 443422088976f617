"""The documents cite evidence by path (profiles/..., tools/..., tests/...): every cited path must exist in the tree, so that a
renamed artefact or a summary that was never written shows up here and not in a review.  CPU only."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOCS = ["DESIGN.md", "README.md", "INTEGRATION.md", "tools/README.md", "tests/README.md", "profiles/r02_summary.md"]
SOURCES = ["bench.py", "__graft_entry__.py"] + [os.path.relpath(p, ROOT) for pat in ("spin-nerf_b200/csrc/*.cu", "spin-nerf_b200/csrc/*.cuh",
                                                                                   "spin-nerf_b200/*.py", "tools/*.py", "tools/*.sh", "include/*.h")
                                                for p in glob.glob(os.path.join(ROOT, pat))]
PATH_RE = re.compile(r"\b((?:profiles|tools|tests|oracle|include)/[A-Za-z0-9_./\-*]+)")
# deliberately absent: the reference is Python, there is nothing to compile into oracle/_ref
ABSENT = ("oracle/_ref",)


def cited_paths(text):
    for m in PATH_RE.finditer(text):
        p = m.group(1).rstrip("./")
        if p.endswith(("/", "_")) or p.count("/") == 0:
            continue
        yield p


def exists(p):
    full = os.path.join(ROOT, p)
    return bool(glob.glob(full)) or bool(glob.glob(full + "*")) if "*" in p else os.path.exists(full)


def test_every_cited_path_exists():
    missing = []
    for doc in DOCS + SOURCES:
        path = os.path.join(ROOT, doc)
        if not os.path.isfile(path):
            continue
        for p in cited_paths(open(path, errors="replace").read()):
            if p.startswith("profiles/") or doc in DOCS:
                if not exists(p) and not p.startswith(ABSENT):
                    missing.append((doc, p))
    assert not missing, missing
