#!/usr/bin/env python
"""Stage the UNMODIFIED reference sources of the hot path for `bench.py --impl reference`.

    python baseline/make_ref.py          (authoring container only: /root/reference does not exist on the GPU box)

Copies models/diffusion.py, models/__init__.py, sdes.py, nets.py, losses.py from /root/reference into the git-ignored
(but gpurun-shipped) directory baseline/_ref/, byte for byte.  Nothing under baseline/_ref is ever committed.  The three
import shims the reference needs (`overrides`, `include.sdeflow_light.lib.utils`, `FrEIA`: modules absent from the
reference tree) are the ones under oracle/shims and stay there.  `__graft_entry__.build()` runs this when
/root/reference is present.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = ["models/__init__.py", "models/diffusion.py", "sdes.py", "nets.py", "losses.py"]


def main():
    if not os.path.isdir(SRC):
        print("make_ref: /root/reference not present; keeping", DST, "as is")
        return 0 if os.path.isdir(DST) else 1
    for f in FILES:
        dst = os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), dst)
    print("make_ref: staged", len(FILES), "reference files in", DST)
    return 0


if __name__ == "__main__":
    sys.exit(main())
