#!/usr/bin/env python
"""Recipe for oracle/_ref: an UNMODIFIED copy of the reference's hot-path Python sources, so that the reference itself
(not only its restatement oracle/mdgan_oracle.py) can be timed on the GPU box's host cores, where /root/reference does
not exist.  oracle/_ref/ is git-ignored (never part of the history) and NOT gpurun-ignored (it travels to the box like a
built .so).  Run by __graft_entry__.build() whenever /root/reference is present; a no-op otherwise.

Copied: src/bootstrap.py, src/standalone_gan.py, src/actors/*.py, src/datasets/*.py (nothing else: no logs, results,
notebooks).  MANIFEST.json records the sha256 of every source file and of its copy (equal by construction; the
manifest is what tests/test_reference_arm.py and bench.py --impl reference check before they run it).

Test / measurement infrastructure only: nothing under distributed-gan_b200/ imports or executes oracle/_ref.
"""
import hashlib
import json
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
DEST = HERE.parent / "_ref"
FILES = ["bootstrap.py", "standalone_gan.py"]
DIRS = ["actors", "datasets"]


def _sha(p: Path) -> str:
    return hashlib.sha256(p.read_bytes()).hexdigest()


def install(ref_root: Path = Path("/root/reference")) -> bool:
    src = ref_root / "src"
    if not src.is_dir():
        return False
    out = DEST / "src"
    if out.exists():
        shutil.rmtree(out)
    out.mkdir(parents=True)
    manifest = {}
    todo = [src / f for f in FILES]
    for d in DIRS:
        todo += sorted((src / d).glob("*.py"))
    for f in todo:
        rel = f.relative_to(src)
        (out / rel).parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(f, out / rel)
        manifest[str(rel)] = {"sha256_source": _sha(f), "sha256_copy": _sha(out / rel)}
    (DEST / "MANIFEST.json").write_text(json.dumps(
        {"reference": "owengombas/distributed-gan", "source_root": str(src), "files": manifest}, indent=1))
    return True


def verify() -> bool:
    """True when oracle/_ref holds every manifest file with the recorded hash."""
    mf = DEST / "MANIFEST.json"
    if not mf.exists():
        return False
    files = json.loads(mf.read_text())["files"]
    return all((DEST / "src" / rel).exists() and _sha(DEST / "src" / rel) == rec["sha256_source"]
               for rel, rec in files.items())


if __name__ == "__main__":
    ok = install()
    print("oracle/_ref installed" if ok else "no /root/reference here: nothing installed", "| verified:", verify())
    sys.exit(0)
