#!/usr/bin/env python3
"""CSV image row (label,p0,p1,...) -> image.ptxt (label,h,w,c,p0,p1,...): the input of encrypt-image.
Same command line and output format as the reference's client/image_converter.py (--format mnist|cifar-10|imagenet, --image PATH)."""
import argparse
import sys

SHAPES = {"mnist": (28, 28, 1), "cifar-10": (32, 32, 3), "imagenet": (224, 224, 3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--format", required=True, type=str.lower, choices=sorted(SHAPES))
    ap.add_argument("--image", required=True)
    ap.add_argument("--row", type=int, default=0, help="row of the CSV to convert (default: the first)")
    ap.add_argument("--out", default="image.ptxt")
    a = ap.parse_args()
    h, w, c = SHAPES[a.format]
    with open(a.image) as f:
        rows = [l for l in f.read().splitlines() if l and l[0].isdigit()]
    fields = [v for v in rows[a.row].split(",") if v != ""]
    if len(fields) != 1 + h * w * c:
        sys.exit(f"{a.image}: {len(fields) - 1} pixels, {a.format} needs {h * w * c}")
    with open(a.out, "w") as f:
        f.write(",".join([fields[0], str(h), str(w), str(c)] + fields[1:]))


if __name__ == "__main__":
    main()
