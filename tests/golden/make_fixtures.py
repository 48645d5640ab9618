"""Regenerates the committed input fixtures from the reference tree (run in the build container only).

  suzanne.npz : the Suzanne mesh of testbed/suzanne_data.h (507 vertices / normals, 968 faces whose
                vertex and normal indices coincide) as float32 / int32 arrays.  It is scene INPUT data;
                nothing under tests/ reads /root/reference at test time.
"""
import re
import sys
from pathlib import Path

import numpy as np

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parent


def block(text, name):
    m = re.search(name + r"\[\]\S*\s*=\s*\{(.*?)\n\};", text, re.S)
    return m.group(1)


def main():
    text = (REF / "testbed" / "suzanne_data.h").read_text()
    num = r"-?\d+\.\d+(?:[eE][-+]?\d+)?"
    verts = np.array(re.findall(num, block(text, "suzanne_vertices")), dtype=np.float64).astype(np.float32).reshape(-1, 3)
    norms = np.array(re.findall(num, block(text, "suzanne_normals")), dtype=np.float64).astype(np.float32).reshape(-1, 3)
    faces = np.array(re.findall(r"-?\d+", block(text, "suzanne_faces")), dtype=np.int32).reshape(-1, 3, 2)
    assert verts.shape == (507, 3) and norms.shape == (507, 3) and faces.shape == (968, 3, 2)
    assert (faces[:, :, 0] == faces[:, :, 1]).all(), "vertex and normal indices are expected to coincide"
    np.savez_compressed(OUT / "suzanne.npz", pos=verts, nrm=norms, faces=np.ascontiguousarray(faces[:, :, 0]))
    print("wrote", OUT / "suzanne.npz", verts.shape, norms.shape, faces.shape)


if __name__ == "__main__":
    main()
