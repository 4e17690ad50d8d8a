"""Collects the reference-derived test fixtures under tests/golden/.

Run in the build container (the only place /root/reference exists):
    python tools/make_golden.py
It copies DATA only (no reference source): the sample image with the reference's
own shipped probability maps (the only known-answer vectors in the tree, SURVEY.md
§4), the two legacy checkpoints that are present in full, and the small
.index/hp/mean/std files of every model folder so the loader and the
shape-derivation code are tested against the real on-disk formats.
"""
import os
import shutil
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SAMPLE = ["UNet sample data/registration/105.tif",
          "UNet sample data/prob_maps/105_ContoursPM_1.tif",
          "UNet sample data/prob_maps/105_NucleiPM_1.tif"]
FULL_MODELS = ["nucleiDAPI", "CytoplasmIncell"]            # .data shard present in the reference tree
ALL_MODELS = ["nucleiDAPI", "nucleiDAPI1-5", "nucleiDAPILAMIN", "CytoplasmIncell2", "CytoplasmIncell",
              "CytoplasmZeissNikon", "mousenucleiDAPI"]
SMALL = ["hp.data", "datasetMean.data", "datasetStDev.data", "model.ckpt.index"]


def main():
    os.makedirs(os.path.join(DST, "sample"), exist_ok=True)
    for rel in SAMPLE:
        shutil.copyfile(os.path.join(REF, rel), os.path.join(DST, "sample", os.path.basename(rel)))
    for m in ALL_MODELS:
        d = os.path.join(DST, "models", m)
        os.makedirs(d, exist_ok=True)
        for f in SMALL:
            shutil.copyfile(os.path.join(REF, "models", m, f), os.path.join(d, f))
        if m in FULL_MODELS:
            f = "model.ckpt.data-00000-of-00001"
            shutil.copyfile(os.path.join(REF, "models", m, f), os.path.join(d, f))
    for root, _, files in os.walk(DST):
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    print("fixtures written to", DST)


if __name__ == "__main__":
    main()
