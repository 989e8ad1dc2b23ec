"""Generates tests/golden/frame_ingest_v1.npz by running the REAL reference frame ingest
(`utils/dataset.py:47-69` from_image_to_embedding_no_IncepRes, exec'd from the reference's source text at run time together with
its transforms_val / constants - importing utils.dataset itself would switch torch's default tensor type) on seeded PNG crops.
Only inputs + outputs are committed. Build container only: python tests/golden/make_ingest_golden.py"""
import os
import re
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF_FILE = "/root/reference/utils/dataset.py"


def reference_ingest():
    import cv2
    import torch
    from PIL import Image
    from torchvision import transforms
    src = open(REF_FILE).read()
    consts = "\n".join(re.findall(r"^(?:NORMAL_MEAN|NORMAL_STD|SWIN_IMG_SIZE)\s*=.*$", src, flags=re.M))
    tv = src[src.index("transforms_val = transforms.Compose("):src.index("def from_image_to_embedding_no_IncepRes")]
    fn = src[src.index("def from_image_to_embedding_no_IncepRes"):src.index("'''加载aff-wild2数据集'''")]
    ns = {"cv2": cv2, "torch": torch, "Image": Image, "transforms": transforms}
    exec(compile(consts + "\n" + tv + "\n" + fn, "reference:utils/dataset.py", "exec"), ns)
    return ns["from_image_to_embedding_no_IncepRes"]


def main():
    import cv2
    fn = reference_ingest()
    rng = np.random.default_rng(1111)

    def smooth(h, w, seed):   # low-frequency content (compresses well; faces are not white noise either)
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
        ch = [127.5 + 127.5 * np.sin(xx / (7.0 + seed + c) + yy / (11.0 + c)) * np.cos(yy / (5.0 + seed)) for c in range(3)]
        return np.clip(np.rint(np.stack(ch, -1)), 0, 255).astype(np.uint8)

    crops = [rng.integers(0, 256, (112, 112, 3), dtype=np.uint8),     # 2x bicubic (BASELINE.json's synthetic case), noise
             smooth(112, 112, 1),                                      # 2x bicubic, smooth
             rng.integers(0, 256, (56, 56, 3), dtype=np.uint8),       # 4x bicubic
             smooth(100, 100, 2),                                      # non-integer bicubic (IPP wheels differ by 1 LSB)
             smooth(224, 224, 3),                                      # no resize
             smooth(448, 448, 4)[:, :, :] // 8 * 8,                    # 2x area (integer ratio); coarse values compress well
             smooth(300, 300, 5) // 8 * 8]                             # general area
    out = {}
    res = {}
    for ipp in (True, False):
        cv2.ipp.setUseIPP(ipp)
        with tempfile.TemporaryDirectory() as td:
            paths = []
            for i, c in enumerate(crops):
                p = os.path.join(td, f"{i}.png")
                cv2.imwrite(p, c)          # lossless; imread returns the same B,G,R bytes
                paths.append(p)
            res[ipp] = fn(paths, "test").astype(np.float32)
    for i, c in enumerate(crops):
        out[f"crop{i}"] = c
    # the 8-bit values behind the float output are enough to store: x = (v/255 - .5)/.5
    u8 = {k: np.rint((v * 0.5 + 0.5) * 255.0) for k, v in res.items()}
    for k in res:
        assert np.abs(((u8[k] / 255.0).astype(np.float32) - 0.5) / 0.5 - res[k]).max() < 1e-6
    import hashlib
    ref = u8[False].astype(np.uint8)                               # OpenCV's own code path
    out["ref_noipp_u8_crop0"] = ref[0]                             # the BASELINE.json case in full, the rest as digests
    out["ref_noipp_sha256"] = np.array([hashlib.sha256(np.ascontiguousarray(ref[i]).tobytes()).hexdigest() for i in range(len(crops))])
    out["ipp_mismatch_fraction"] = np.array([(u8[True][i] != u8[False][i]).mean() for i in range(len(crops))])
    out["ipp_max_abs_diff"] = np.array([np.abs(u8[True][i] - u8[False][i]).max() for i in range(len(crops))])
    out["ref_f32_samples"] = res[False][:, :, ::97, ::89].copy()   # pins the normalisation arithmetic bit for bit
    print("IPP vs native mismatch fraction per crop:", np.round(out["ipp_mismatch_fraction"], 4), out["ipp_max_abs_diff"])
    p = os.path.join(ROOT, "tests", "golden", "frame_ingest_v1.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, os.path.getsize(p), "bytes")


if __name__ == "__main__":
    main()
