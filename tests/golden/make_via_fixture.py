"""Realistic VIA outlines for tests/test_via_polygons.py: the first polygons of the annotation files the reference ships
(datasets/{rice,food}/{train,val}/via_*_annotation.json; the images themselves are not in the repository), written as
tests/golden/via_polygons_fixture.json.  Inputs only -- the reference holds no rasterised masks to compare with, and
scikit-image is not installable here.  Run in the build container: python tests/golden/make_via_fixture.py"""
import json
import os

REF = "/root/reference/datasets"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "via_polygons_fixture.json")


def main():
    out = []
    for ds, sub, name, take in (("rice", "train", "via_rice_annotation.json", 10), ("rice", "val", "via_rice_annotation.json", 3),
                                ("food", "train", "via_food_annotation.json", 10), ("food", "val", "via_food_annotation.json", 3)):
        ann = json.load(open(os.path.join(REF, ds, sub, name)))
        n = 0
        for a in ann.values():
            regions = a["regions"]
            regions = list(regions.values()) if isinstance(regions, dict) else regions
            polys = [r["shape_attributes"] for r in regions if r["shape_attributes"].get("name") == "polygon"]
            if not polys:
                continue
            out.append({"source": "%s/%s/%s" % (ds, sub, a["filename"]),
                        "polygons": [{"all_points_x": p["all_points_x"], "all_points_y": p["all_points_y"]} for p in polys]})
            n += 1
            if n >= take:
                break
    json.dump(out, open(OUT, "w"), separators=(",", ":"))
    print(len(out), "images,", sum(len(o["polygons"]) for o in out), "polygons ->", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
