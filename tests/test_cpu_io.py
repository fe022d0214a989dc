"""Asset readers either side of the path: the OpenEXR reader incl. the PIZ decoder
(hm_piz.cpp) against fixtures written and read back by OpenCV's OpenEXR (tests/golden/gen_piz_fixture.py),
and — where the reference tree is mounted — the shipped scenes themselves."""
import os

import numpy as np
import pytest

from hairmsnn_b200 import api

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = "/root/reference/scenes"


@pytest.mark.parametrize("name", ["piz_f32", "piz_f16"])
def test_piz_fixture_decodes_bit_exact(name):
    img = api.load_exr(os.path.join(GOLD, name + ".exr"))
    want = np.load(os.path.join(GOLD, name + ".npy"))
    assert img.shape == (77, 131, 4)
    assert np.array_equal(img[..., :3].view(np.uint32), want.view(np.uint32))
    assert np.all(img[..., 3] == 1.0)


def test_corrupt_piz_block_is_an_error_not_a_crash(tmp_path):
    raw = bytearray(open(os.path.join(GOLD, "piz_f32.exr"), "rb").read())
    for k in range(len(raw) - 4000, len(raw) - 3000):
        raw[k] ^= 0x5A
    p = tmp_path / "bad.exr"
    p.write_bytes(bytes(raw))
    try:
        img = api.load_exr(str(p))          # damage may decode to wrong pixels, but must not crash
        assert img.shape == (77, 131, 4)
    except api.HairMSNNError as e:
        assert e.code == -1
    p2 = tmp_path / "short.exr"
    p2.write_bytes(bytes(raw[:5000]))
    with pytest.raises(api.HairMSNNError):
        api.load_exr(str(p2))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference scenes are not mounted here")
def test_shipped_straight_scene_loads_unchanged():
    """scenes/straight/config.json as shipped: Windows absolute paths, wrong-case .hair name, PIZ env map."""
    sc = api.Scene.load(os.path.join(REF, "straight", "config.json"))
    i = sc.info()
    assert (i.num_segments, i.num_strands, i.num_triangles) == (1200000, 50000, 78520)
    assert (i.width, i.height, i.spp, i.path_v1, i.path_v2) == (1024, 1024, 500, 1, 40)
    assert (i.env_w, i.env_h, i.num_dlights) == (4096, 2048, 1)
    assert 170 < i.scene_scale < 182                      # ~176
    env = sc.env_tables()["env"].reshape(2048, 4096, 4)
    cv2 = pytest.importorskip("cv2")
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    ref = cv2.imread(os.path.join(REF, "envmaps", "christmas_photo_studio_07_4k.exr"), cv2.IMREAD_UNCHANGED)
    if ref is None:
        pytest.skip("OpenCV build has no OpenEXR")
    assert np.array_equal(env[..., :3], ref[..., ::-1])


def test_bvh_cache_roundtrip(tmp_path, monkeypatch):
    """HM_BVH_CACHE: the second construction of the same geometry loads the wide tree instead of building."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from common import small_scene_kwargs
    kw = small_scene_kwargs(strands=200, segs=10)
    a = api.Scene.from_arrays(**kw)                       # no cache
    monkeypatch.setenv("HM_BVH_CACHE", str(tmp_path))
    b = api.Scene.from_arrays(**kw)                       # builds and stores
    files = [f for f in os.listdir(tmp_path) if f.startswith("hm_bvh_") and f.endswith(".bin")]
    assert len(files) == 1
    c = api.Scene.from_arrays(**kw)                       # restored
    ia, ib, ic = a.info(), b.info(), c.info()
    assert ib.num_bvh_nodes == ia.num_bvh_nodes > 0 and ic.num_bvh_nodes == 0
    assert (ic.num_wide_nodes, ic.num_wide_leaf_refs, ic.wide_depth) == (ia.num_wide_nodes, ia.num_wide_leaf_refs, ia.wide_depth)
    assert ia.num_wide_nodes > 0 and ia.wide_depth >= 2
    # other geometry -> other key
    kw2 = small_scene_kwargs(strands=201, segs=10)
    api.Scene.from_arrays(**kw2)
    assert len([f for f in os.listdir(tmp_path) if f.endswith(".bin")]) == 2
    # a damaged file is ignored and rebuilt over
    path = os.path.join(tmp_path, files[0])
    open(path, "r+b").write(b"garbage!")
    d = api.Scene.from_arrays(**kw)
    assert d.info().num_bvh_nodes == ia.num_bvh_nodes


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference scenes are not mounted here")
@pytest.mark.parametrize("name", ["studio_small_01_4k.exr", "christmas_photo_studio_07_4k.exr"])
def test_shipped_env_maps_decode_bit_exact(name):
    cv2 = pytest.importorskip("cv2")
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    path = os.path.join(REF, "envmaps", name)
    ref = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if ref is None:
        pytest.skip("OpenCV build has no OpenEXR")
    img = api.load_exr(path)
    assert img.shape == (2048, 4096, 4)
    assert np.array_equal(img[..., :3], ref[..., ::-1]) and np.all(img[..., 3] == 1.0)


def _scene_dir():
    for d in (REF, os.path.join(os.path.dirname(GOLD), "..", "assets", "scenes")):
        if os.path.isdir(d):
            return os.path.abspath(d)
    return None


@pytest.mark.skipif(_scene_dir() is None or not os.path.exists(os.path.join(os.path.dirname(GOLD), "..", "oracle", "_ref", "libref_scene.so")),
                    reason="needs the shipped scenes and oracle/_ref/libref_scene.so (the reference's scene.cpp compiled for the host)")
@pytest.mark.parametrize("scene,hair", [("curly", "wCurly.hair"), ("straight", "wStraight.hair")])
def test_hair_reader_matches_reference_extract_hair_data(scene, hair):
    """Scene::extractHairData (scene.cpp:10-73, the REFERENCE's own source compiled for the host) vs
    the product's .hair reader, array by array, for both shipped files: control points incl. the mirrored phantom end
    points, radii (0.2 x thickness), first-control-point index of every segment, bounds grown from the origin."""
    import ctypes as C
    path = os.path.join(_scene_dir(), scene, hair)
    lib = C.CDLL(os.path.join(os.path.dirname(GOLD), "..", "oracle", "_ref", "libref_scene.so"))
    counts = (C.c_int * 3)()
    assert lib.ref_extract_hair(os.fsencode(path), counts) == 0
    ncp, nseg, nstrands = counts[0], counts[1], counts[2]
    cps = np.zeros((ncp, 3), np.float32); w = np.zeros(ncp, np.float32); seg = np.zeros(nseg, np.int32)
    b = np.zeros(6, np.float32); scale = C.c_float()
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    lib.ref_hair_arrays(cps.ctypes.data_as(fp), w.ctypes.data_as(fp), seg.ctypes.data_as(ip), b.ctypes.data_as(fp), C.byref(scale))
    got = api.load_hair_file(path)
    want_counts = {"curly": (3541580, 3391580, 50000), "straight": (1350000, 1200000, 50000)}[scene]
    assert (ncp, nseg, nstrands) == want_counts
    assert (got["cps"].shape[0], got["seg_cp"].shape[0], got["strands"]) == want_counts
    assert np.array_equal(got["cps"][:, :3].view(np.uint32), cps.view(np.uint32))
    assert np.array_equal(got["cps"][:, 3].view(np.uint32), w.view(np.uint32))
    assert np.array_equal(got["seg_cp"], seg)
    assert np.array_equal(got["bounds"][0], b[:3]) and np.array_equal(got["bounds"][1], b[3:])
    d = got["bounds"][1] - got["bounds"][0]
    assert abs(float(np.sqrt((d * d).sum(dtype=np.float32))) - scale.value) < 1e-3
