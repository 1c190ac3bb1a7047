"""Recipe for `oracle/_ref/`: the UNMODIFIED reference sources of the hot path, placed where the GPU box can import them.

    python oracle/make_ref.py          (build container only: needs /root/reference)

The reference is pure Python (no build step): the files its TEDM inference path imports are copied byte for byte from
where they lie under /root/reference into `oracle/_ref/` (git-ignored -- reference sources never enter this repository's
history -- but shipped to the GPU box with the snapshot, like the built .so).  `MANIFEST.json` records the SHA-256 of
every file so a test can show the copy is unmodified.  Test infrastructure: only bench.py's CPU legs (`cpu_baseline`,
`--impl reference`) and tests/ import it, as the thing that is timed BESIDE the product, never as part of it.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("TEDM_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
# what `DatasetDM(args)(x)` + the shared-weight head + the ensemble need (auxiliary/postprocessing/testing_shared_weights.py:104-144)
FILES = ["models/unet_model.py", "models/diffusion_model.py", "models/datasetDM_model.py", "trainers/utils.py", "LICENSE"]


def sha256(path: str) -> str:
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main() -> int:
    if not os.path.isdir(SRC):
        print(f"{SRC} not present: oracle/_ref left as it is")
        return 0
    manifest = {}
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        manifest[rel] = sha256(dst)
    json.dump({"source": "mmr12/TEDM (read-only copy at /root/reference)", "sha256": manifest},
              open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    print(f"oracle/_ref: {len(FILES)} files copied from {SRC}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
