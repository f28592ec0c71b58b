"""Recipe for ``oracle/_ref`` — the reference's own implementation of the hot path, for the CPU baseline.

ORACLE side (test / baseline infrastructure).  The reference is pure Python: nothing to compile.  This script copies
exactly the two files that hold the hot path, UNMODIFIED, from where they lie under ``/root/reference``:

    yolox_24p/models/losses.py     IOUloss, Loss_Function (SimOTA, loss)            SURVEY.md 8(a) rows a1-a8
    yolox_24p/utils/boxes.py       postprocess, circle_inter, bboxes_iou            rows a5, a9

into ``oracle/_ref/yolox_24p/`` (git-ignored: reference sources never enter the history; NOT gpurun-ignored: the copy
travels to the GPU box like a built ``.so``) and writes two package ``__init__`` shims of its own (the reference's
``__init__`` files import the CNN, datasets and loggers, which the hot path does not need).  A manifest with the
SHA-256 of both files is written next to them; ``oracle/ref_runtime.py`` loads the tree.

    python oracle/make_ref.py [--reference /root/reference]
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref", "yolox_24p")
FILES = ["models/losses.py", "utils/boxes.py"]
SHIMS = {
    "models/__init__.py": "from .losses import IOUloss, Loss_Function  # shim written by oracle/make_ref.py\n",
    "utils/__init__.py": "from .boxes import *  # shim written by oracle/make_ref.py\n"
                         "from .boxes import bboxes_iou, circle_inter, postprocess\n",
}


def make(reference: str = "/root/reference") -> bool:
    src_root = os.path.join(reference, "yolox_24p")
    if not all(os.path.isfile(os.path.join(src_root, f)) for f in FILES):
        return False
    manifest = {}
    for f in FILES:
        dst = os.path.join(DEST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src_root, f), dst)
        with open(dst, "rb") as fh:
            manifest[f] = hashlib.sha256(fh.read()).hexdigest()
    for f, text in SHIMS.items():
        with open(os.path.join(DEST, f), "w") as fh:
            fh.write(text)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src_root, "sha256": manifest}, fh, indent=1)
    return True


if __name__ == "__main__":
    ref = sys.argv[sys.argv.index("--reference") + 1] if "--reference" in sys.argv else "/root/reference"
    ok = make(ref)
    print("oracle/_ref written" if ok else f"no reference tree under {ref}: nothing to do")
