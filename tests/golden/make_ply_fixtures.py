"""Regenerates tests/golden/ply_fixtures.json from the reference's own PLY test files (src/test_data/*.ply) and the
values its unit tests assert on them (src/read_write/ply.rs:746-790).  Needs /root/reference, so it runs only in the
development container; the JSON travels with the repo.

    python tests/golden/make_ply_fixtures.py
"""
import json
import os

REF = "/root/reference/src/test_data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ply_fixtures.json")

# ply.rs tests: BATCH_SIZE = 2 -> NUM_BATCHES = 4; first position x == 1, last position x == 22;
# first colour red == 255; last colour red as listed; the intensity file carries 8 NaN intensities.
EXPECT = {
    "xyz_f32_rgb_u8_le.ply": dict(num_points=8, batches_of_2=4, first_x=1.0, last_x=22.0, first_red=255, last_red=234, intensity=False),
    "xyz_f32_rgba_u8_le.ply": dict(num_points=8, batches_of_2=4, first_x=1.0, last_x=22.0, first_red=255, last_red=227, intensity=False),
    "xyz_f32_rgb_u8_intensity_f32.ply": dict(num_points=8, batches_of_2=4, first_x=1.0, last_x=22.0, first_red=255, last_red=234, intensity=True,
                                             intensity_all_nan=True),
}

if __name__ == "__main__":
    out = {}
    for name, exp in EXPECT.items():
        with open(os.path.join(REF, name), "rb") as f:
            out[name] = dict(hex=f.read().hex(), expect=exp)
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT, {k: len(v["hex"]) // 2 for k, v in out.items()})
