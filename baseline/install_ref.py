"""Copies the UNMODIFIED reference sources of the hot path (drqv2.py, utils.py, replay_buffer.py) from
/root/reference into baseline/_ref/ (git-ignored, but it travels to the GPU box with the snapshot), so
that `bench.py --impl reference` can time the reference's own CPU implementation there.  The
reference is pure Python (nothing to compile); hydra / omegaconf, which it imports but never uses on
this path, are stubbed at import time by the caller.  Run by __graft_entry__.build()
whenever /root/reference exists."""
import pathlib
import shutil
import sys

SRC = pathlib.Path("/root/reference")
DST = pathlib.Path(__file__).resolve().parent / "_ref"


def install():
    if not SRC.exists():
        return False
    DST.mkdir(parents=True, exist_ok=True)
    for name in ("drqv2.py", "utils.py", "replay_buffer.py"):
        dst = DST / name
        if dst.exists():
            dst.chmod(0o644)
        shutil.copyfile(SRC / name, dst)
        dst.chmod(0o444)
    return True


if __name__ == "__main__":
    print("installed" if install() else "reference not present", file=sys.stderr)
