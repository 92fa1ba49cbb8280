#!/usr/bin/env python
"""Regenerate mdir_b200/data/rgb2lab_lut_s16.npy: the 33^3 lattice of OpenCV's float
RGB->Lab conversion (cv2.cvtColor(float32, COLOR_RGB2LAB) does NOT evaluate the Lab formulas per
pixel: it trilinearly interpolates a fixed-point table, color_lab.cpp RGB2Lab_f / trilinearInterpolate).
Feeding cv2 the lattice points themselves returns the table entries exactly, so the table is
recovered from the installed wheel (opencv-python is the arbiter the reference names,
requirements.txt:3).  Channels: round(16384*L/100), round(16384*(a+128)/256), round(16384*(b+128)/256)."""
import os

import cv2
import numpy as np

g = (np.arange(33) / 32.0).astype(np.float32)
R, G, B = np.meshgrid(g, g, g, indexing="ij")
lab = cv2.cvtColor(np.stack([R, G, B], -1).reshape(-1, 1, 3), cv2.COLOR_RGB2LAB).reshape(33, 33, 33, 3).astype(np.float64)
lut = np.stack([np.rint(lab[..., 0] / 100 * 16384), np.rint((lab[..., 1] + 128) / 256 * 16384), np.rint((lab[..., 2] + 128) / 256 * 16384)], -1)
assert lut.min() >= 0 and lut.max() <= 32767
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mdir_b200", "data", "rgb2lab_lut_s16.npy")
np.save(out, lut.astype(np.int16))
print("wrote", out, lut.shape, "cv2", cv2.__version__)
