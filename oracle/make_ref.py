"""
TEST INFRASTRUCTURE — recipe for ``oracle/_ref/``: the UNMODIFIED reference package, placed next
to the oracle so that it travels to the GPU box with the repository snapshot (``oracle/_ref/`` is
git-ignored, not gpurun-ignored; /root/reference itself does not exist there).

    python oracle/make_ref.py          # /root/reference/napkon_string_matching -> oracle/_ref/

The reference is pure Python (no build step): the recipe copies the package's ``.py`` files
byte for byte (its own tests and READMEs are left out) and writes ``MANIFEST.json`` with the
sha256 of every file, so that a run can state exactly which sources it timed.  Nothing under
``oracle/_ref/`` is ever committed or imported by the product; ``oracle/ref_arm.py`` imports it in
child processes (with ``oracle/shims`` standing in for the absent ``nltk`` / ``rapidfuzz`` /
``psycopg2``) as the CPU arm of ``bench.py`` and to validate the oracle port.
"""
from __future__ import annotations

import hashlib
import json
import pathlib
import shutil
import sys

HERE = pathlib.Path(__file__).resolve().parent
REFERENCE = pathlib.Path("/root/reference")
DEST = HERE / "_ref"
PACKAGE = "napkon_string_matching"


def make(reference: pathlib.Path = REFERENCE, dest: pathlib.Path = DEST) -> pathlib.Path | None:
    """Returns ``dest`` (rebuilt) or None when the reference tree is not on this machine."""
    src = reference / PACKAGE
    if not src.is_dir():
        return None
    out = dest / PACKAGE
    if out.exists():
        shutil.rmtree(out)
    manifest = {}
    for path in sorted(src.rglob("*.py")):
        rel = path.relative_to(src)
        if rel.parts[0] == "tests":
            continue
        target = out / rel
        target.parent.mkdir(parents=True, exist_ok=True)
        data = path.read_bytes()
        target.write_bytes(data)
        manifest[str(rel)] = hashlib.sha256(data).hexdigest()
    (dest / "MANIFEST.json").write_text(json.dumps(
        {"source": str(src), "files": manifest,
         "note": "unmodified copies; git-ignored; test infrastructure only"}, indent=1))
    return dest


def available(dest: pathlib.Path = DEST) -> bool:
    return (dest / PACKAGE / "types" / "comparable_data.py").exists()


if __name__ == "__main__":
    made = make()
    print(f"built {made}" if made else f"{REFERENCE} not present: nothing to do", file=sys.stderr)
