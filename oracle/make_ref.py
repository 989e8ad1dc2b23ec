"""TEST / BASELINE INFRASTRUCTURE ONLY -- recipe that makes the real reference travel to the GPU box.

`/root/reference` (NUSTM/FacialMMT, pure Python) exists only in the build container. `bench.py --impl reference` and the
`cpu_baseline` leg time the reference's OWN modules on the GPU box's host cores, so the files of the path are copied,
byte for byte and unmodified, into the git-ignored `baseline/_ref/` (listed in .gitignore, NOT in .gpurunignore: it ships
with the gpurun snapshot like the built .so, and never enters the history). Nothing under `baseline/_ref/` is imported
by the product package; only oracle/ref_harness.py puts it on sys.path, for tests/, smoke() and bench.py's CPU legs.

    python oracle/make_ref.py            # no-op (exit 0) where /root/reference is absent

`__graft_entry__.build()` runs this in the build container.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("FMMT_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")

# the files of the hot path + the eval loop and metric that call it
FILES = [
    "src/models.py",
    "modules/Transformer.py",
    "modules/CrossmodalTransformer.py",
    "modules/multihead_attention.py",
    "modules/position_embedding.py",
    "modules/SwinTransformer/Swin_Transformer.py",
    "modules/SwinTransformer/backbone_def.py",
    "modules/SwinTransformer/swin_conf.yaml",
    "train.py",
    "utils/eval_metrics.py",
    "utils/dataset.py",            # from_image_to_embedding_no_IncepRes (frame ingest), exec'd from its source text
    "src/meld_bert_extraText.py",
    "LICENSE",
]


def make(verbose: bool = True) -> bool:
    if not os.path.isfile(os.path.join(SRC, "src", "models.py")):
        if verbose:
            print(f"[make_ref] {SRC} not present: nothing to do (the GPU box uses the prebuilt baseline/_ref)")
        return False
    lines = []
    for rel in FILES:
        s = os.path.join(SRC, rel)
        if not os.path.isfile(s):
            continue
        d = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        lines.append(f"{hashlib.sha256(open(d, 'rb').read()).hexdigest()}  {rel}")
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(lines) + "\n")
    if verbose:
        print(f"[make_ref] copied {len(lines)} reference files (unmodified) into {DST}")
    return True


if __name__ == "__main__":
    make()
    sys.exit(0)
