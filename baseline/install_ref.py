"""Place an UNMODIFIED copy of the reference's Python tree under ``baseline/_ref/`` so it can travel to the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY.  ``baseline/_ref/`` is git-ignored (never part of this repository's history)
but not gpurun-ignored, exactly like the compiled reference kernel in ``oracle/_ref/``.  The reference is a plain
script tree (no setup.py for the Python part), so "installing" it is a file copy of its .py / .gin files; nothing is
edited.  ``baseline/refrun.py`` imports it from there through the three import shims in ``oracle/shims``.

    python baseline/install_ref.py          # run in the build container (needs /root/reference)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CER_REFERENCE_DIR", "/root/reference")
DST = os.path.join(HERE, "_ref")
KEEP_EXT = (".py", ".gin", ".md", ".yml", ".cpp", ".cu")


def installed() -> bool:
    return os.path.isfile(os.path.join(DST, "core", "raft.py"))


def install(force: bool = False) -> str:
    if not os.path.isfile(os.path.join(REF, "core", "raft.py")):
        if installed():
            return DST
        raise RuntimeError("reference tree not present and baseline/_ref not installed")
    if installed() and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for root, dirs, files in os.walk(REF):
        dirs[:] = [d for d in dirs if not d.startswith(".")]
        rel = os.path.relpath(root, REF)
        for f in files:
            if f.endswith(KEEP_EXT) or f == "LICENSE":
                os.makedirs(os.path.join(DST, rel), exist_ok=True)
                shutil.copyfile(os.path.join(root, f), os.path.join(DST, rel, f))
    for root, dirs, files in os.walk(DST):          # the copy is writable (the source tree is read-only)
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    return DST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
