"""Generates the PIZ-compressed EXR fixtures for tests/test_cpu_io.py with OpenCV's OpenEXR writer
(an independent implementation of the format): run `OPENCV_IO_ENABLE_OPENEXR=1 python gen_piz_fixture.py`.
Committed outputs: piz_f32.exr, piz_f16.exr (PIZ), and the expected pixels piz_f32.npy, piz_f16.npy
as OpenCV reads them back."""
import os
os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2
import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(20231017)
H, W = 77, 131      # odd sizes, more than two 32-line PIZ blocks, last block partial
y, x = np.mgrid[0:H, 0:W].astype(np.float32)
img = np.stack([np.sin(x / 9) * np.cos(y / 7) + 1.5, np.exp(-((x - 60) ** 2 + (y - 30) ** 2) / 300) * 40, (x + y) / 50], axis=2).astype(np.float32)
img += rng.random(img.shape).astype(np.float32) * 0.05
img[10:20, 10:40] = 0.0          # constant run -> run-length escapes
img[50, 100] = 6.0e4             # large dynamic range
for name, typ in (("piz_f32", cv2.IMWRITE_EXR_TYPE_FLOAT), ("piz_f16", cv2.IMWRITE_EXR_TYPE_HALF)):
    path = os.path.join(here, name + ".exr")
    ok = cv2.imwrite(path, img[..., ::-1], [cv2.IMWRITE_EXR_TYPE, typ, cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_PIZ])
    assert ok
    back = cv2.imread(path, cv2.IMREAD_UNCHANGED)[..., ::-1]
    np.save(os.path.join(here, name + ".npy"), np.ascontiguousarray(back.astype(np.float32)))
    print(name, os.path.getsize(path), "bytes")
