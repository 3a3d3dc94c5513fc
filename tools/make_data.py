"""Packs the reference's shipped DATA fixtures (not code) needed at run time on the GPU box, where
/root/reference does not exist:  estimation/landmarks/intrinsics.csv (7 rows) and the 34 landmark CSVs
sim/landmark_csvs/<MGRS>_top_salient.csv (16,825 rows x 6 floats) -> vinsat_b200/data/.
Run in the build container:  python tools/make_data.py"""
import csv
import glob
import os
import shutil

import numpy as np

REF = os.environ.get("VINSAT_REF", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vinsat_b200", "data")
os.makedirs(OUT, exist_ok=True)
shutil.copyfile(os.path.join(REF, "estimation", "landmarks", "intrinsics.csv"), os.path.join(OUT, "intrinsics.csv"))
regions, rows, offsets = [], [], [0]
for path in sorted(glob.glob(os.path.join(REF, "sim", "landmark_csvs", "*_top_salient.csv"))):
    regions.append(os.path.basename(path).split("_")[0])
    with open(path) as f:
        r = csv.reader(f)
        next(r)
        data = [[float(x) for x in row] for row in r]
    rows.extend(data)
    offsets.append(len(rows))
np.savez_compressed(os.path.join(OUT, "landmarks_mgrs.npz"), regions=np.array(regions), offsets=np.array(offsets, dtype=np.int64),
                    rows=np.array(rows, dtype=np.float64),
                    columns=np.array(["centroid_lon", "centroid_lat", "tl_lon", "tl_lat", "br_lon", "br_lat"]))
print(len(regions), "regions", len(rows), "landmarks")
