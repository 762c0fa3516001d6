#!/usr/bin/env python
"""Builds tests/golden/reference_images.json from the three renders the reference ships
(/root/reference/images/book1.png, book2.png, stanford_dragon.png; README.md:6,18,34).

These PNGs are the only outputs of the hot path that the reference itself holds (it has no tests beyond vec3 and
draws from an unseeded thread_rng), so they are what pins the oracle and the CUDA path to the reference: the
fixture keeps, per image, its size, the rows that are exactly black (the H - threads * (H / threads) unrendered
rows of world.rs:1198-1202), and per named region the mean LINEAR radiance.  A PNG value v is
(255.9 * sqrt(c)) as i32 (vec3.rs:89-107), so c ~ ((v + 0.5) / 255.9)^2.  Means are compared in linear space
because the mean of sqrt(noisy) is biased low at the few hundred spp a test can afford.

Regions are fractions of the image (x0, x1, y0, y1; y = 0 is the TOP row of the PNG) and were picked on parts
of each picture that the scene CODE fixes (light, fog, walls, the large spheres of final_scene, the sky), not on
parts that depend on the reference's unseeded random placement.  /root/reference is not present on the GPU box:
tests read only the committed JSON.  Run here:  python tools/make_reference_image_fixture.py
"""
import json
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/images"

# name -> (x0, x1, y0, y1, kind): kind "mean" = mean linear radiance, "median" = per-channel median (robust against the
# randomly placed small spheres of book-1), "saturated" = every pixel is 255
REGIONS = {
    "book2": {  # final_scene world.rs:494-616, camera world.rs:1009-1029, 1000x1000, 11 threads
        "light": (0.20, 0.45, 0.04, 0.10, "saturated"),          # XzRect(123,432,147,412,554) (7,7,7)
        "fog_background": (0.70, 0.95, 0.05, 0.25, "mean"),      # ConstantMedium 1e-4 in the r=5000 shell, lit by the light
        "moving_sphere": (0.10, 0.18, 0.29, 0.36, "mean"),       # MovingSphere (400,400,400)->(430,..) Lambertian(0.7,0.3,1)
        "blue_medium_sphere": (0.22, 0.32, 0.70, 0.80, "mean"),  # glass r=70 + medium (0.2,0.4,0.9) density 0.2
        "noise_sphere": (0.40, 0.48, 0.45, 0.54, "mean"),        # Noise(0.1) r=80 (tables random; mean tone only)
        "sphere_cluster": (0.56, 0.74, 0.33, 0.48, "mean"),      # 1000 r=10 spheres U[0,165)^3, rotated + translated
        "ground_boxes": (0.65, 0.95, 0.85, 0.99, "mean"),        # 20x20 boxes, heights random: mean tone only
        "glass_sphere": (0.45, 0.56, 0.70, 0.82, "mean"),        # Dielectric(1.5) (260,150,45) r=50, refracts random boxes
        "metal_sphere": (0.81, 0.89, 0.68, 0.74, "mean"),        # Metal((0.8,0.8,0.9),1.0) (0,150,145) r=50, reflects random boxes
    },
    "stanford_dragon": {  # room of world.rs:681-747 (mesh region excluded), camera world.rs:1114-1134, 600x375
        "left_wall_green": (0.02, 0.10, 0.05, 0.40, "mean"),
        "right_wall_blue": (0.92, 0.99, 0.05, 0.40, "mean"),
        "backdrop_pink": (0.15, 0.22, 0.05, 0.40, "mean"),
        "mirror_floor": (0.20, 0.30, 0.60, 0.80, "mean"),
        "below_floor_left": (0.02, 0.12, 0.70, 0.95, "mean"),
        "below_floor_right": (0.90, 0.98, 0.70, 0.95, "mean"),
    },
    "book1": {  # book-1 final with the gradient sky (earlier revision of gen_random_scene; random spheres differ)
        "sky_top": (0.01, 0.25, 0.00, 0.05, "mean"),
        "sky_upper_right": (0.80, 0.99, 0.02, 0.12, "mean"),
        "sky_above_horizon": (0.01, 0.20, 0.15, 0.20, "mean"),
        "ground_far": (0.05, 0.95, 0.26, 0.30, "median"),
        "ground_near": (0.05, 0.95, 0.80, 0.95, "median"),
        "big_lambertian": (0.325, 0.355, 0.14, 0.30, "median"),   # (-4,1,0) r=1 Lambertian(0.4,0.2,0.1): the part left of the glass sphere
        "big_metal_top": (0.60, 0.76, 0.16, 0.28, "median"),      # (4,1,0) r=1 Metal((0.7,0.6,0.5),0): upper half mirrors the sky
        "big_glass_low": (0.42, 0.47, 0.36, 0.44, "median"),      # (0,1,0) r=1 Dielectric(1.5): lower half shows the refracted sky
    },
}


def linear(img_u8):
    return ((img_u8.astype(np.float64) + 0.5) / 255.9) ** 2


def region_pixels(a, r):
    h, w = a.shape[:2]
    x0, x1, y0, y1 = r[:4]
    return a[int(y0 * h):int(y1 * h), int(x0 * w):int(x1 * w)].reshape(-1, a.shape[2])


def region_stat(lin, u8, r):
    kind = r[4]
    px = region_pixels(lin, r)
    if kind == "saturated":
        return {"kind": kind, "all_255": bool((region_pixels(u8, r) == 255).all()), "n": int(px.shape[0])}
    val = px.mean(axis=0) if kind == "mean" else np.median(px, axis=0)
    # saturated = the region touches 255 somewhere: the test then clamps its own pixels at 1 too before averaging
    return {"kind": kind, "linear_rgb": [float(x) for x in val], "n": int(px.shape[0]), "saturated": bool((region_pixels(u8, r) == 255).any())}


def main():
    out = {"_about": "derived from /root/reference/images/*.png by tools/make_reference_image_fixture.py; linear = ((v+0.5)/255.9)^2"}
    for name, regions in REGIONS.items():
        u8 = np.asarray(Image.open(os.path.join(REF, name + ".png")).convert("RGB"))
        lin = linear(u8)
        h, w = u8.shape[:2]
        black = [int(i) for i in range(h) if u8[i].max() == 0]
        rec = {"width": w, "height": h, "black_rows_from_top": black, "regions": {}}
        for rn, r in regions.items():
            rec["regions"][rn] = {"box": list(r[:4]), **region_stat(lin, u8, r)}
        if name == "book1":  # the sky is analytic (primary-ray misses): keep a column of 8-bit values for an exact check
            rec["sky_column_x"] = [4, 8]
            rec["sky_column_rows"] = list(range(0, 100, 4))
            rec["sky_column_u8"] = [[int(round(float(u8[y, 4:8, c].mean()))) for c in range(3)] for y in rec["sky_column_rows"]]
        out[name] = rec
    path = os.path.join(ROOT, "tests", "golden", "reference_images.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)
    return 0


if __name__ == "__main__":
    sys.exit(main())
