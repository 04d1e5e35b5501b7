"""Generate tests/golden/* by running the REFERENCE ITSELF (imported from /root/reference) in this container.

TEST INFRASTRUCTURE ONLY.  Run from the repo root:  python oracle/gen_golden.py
The reference cannot travel to the GPU box, so its outputs on fixed inputs are committed as fixtures, together
with this script.  Nothing of the reference's source is copied: the script imports lib.ImageOperations /
lib.Helpers (cwd = /root/reference because of their relative ./jsons paths) and records what they return.

The only substitution: lib.ImageOperations.fast_cuda_blur (a numba-CUDA kernel, needs a GPU) is replaced by the
integer restatement oracle.restate.blur5_floor; `blur_sim_check` below proves the two bit-identical by running the
reference kernel under NUMBA_ENABLE_CUDASIM=1 on a small random image (recorded in meta.json).
"""
import contextlib
import copy
import hashlib
import io
import json
import os
import subprocess
import sys
import zlib

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)

from oracle import restate as R          # noqa: E402
from mocapv2_b200 import synth as S      # noqa: E402


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def blur_sim_check():
    code = r"""
import os, sys, numpy as np
os.environ['NUMBA_ENABLE_CUDASIM']='1'
os.chdir('/root/reference'); sys.path.insert(0,'/root/reference'); sys.path.insert(0, %r)
from lib.CudaOperations import fast_cuda_blur
from oracle.restate import blur5_floor
rng=np.random.default_rng(7)
ok=True
for shape in [(24,40),(17,9)]:
    img=rng.integers(0,256,shape).astype(np.uint8)
    ok = ok and np.array_equal(fast_cuda_blur(img,5), blur5_floor(img))
print('EQUAL' if ok else 'DIFF')
""" % REPO
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=3000)
    return out.stdout.strip().endswith("EQUAL")


def kat_bundle_adjustment(Hh, ip):
    """Second golden vector of the reference: jsons/before_ba_extrinsics.json --bundle_adjustment--> after_ba_extrinsics.json
    (CalculateCameraPoses.py:243-254).  Stored: the input poses, the reference's result run here, and the shipped JSON."""
    before = json.load(open("jsons/before_ba_extrinsics.json"))
    after = json.load(open("jsons/after_ba_extrinsics.json"))
    poses = [{"R": np.array(p["R"]), "t": np.array(p["t"])} for p in before]
    Hh.camera_params = None
    with quiet():
        out = Hh.bundle_adjustment(ip, poses)
    np.savez_compressed(os.path.join(OUT, "kat_bundle_adjustment.npz"), image_points=ip,
                        R_before=np.stack([p["R"] for p in poses]), t_before=np.stack([p["t"] for p in poses]),
                        R_ref=np.stack([np.asarray(p["R"], dtype=np.float64) for p in out]),
                        t_ref=np.stack([np.asarray(p["t"], dtype=np.float64).ravel() for p in out]),
                        R_json=np.stack([np.array(p["R"]) for p in after]), t_json=np.stack([np.array(p["t"]) for p in after]))
    print("BA max |ref - json| R, t =", np.abs(np.asarray(out[1]["R"]) - np.array(after[1]["R"])).max(),
          np.abs(np.asarray(out[1]["t"]).ravel() - np.array(after[1]["t"])).max())


def kat_fundamentals():
    """Golden vectors of poses_to_fundamental_matrix (CalculateCameraPoses.py:26-78).  That script cannot be imported (it
    opens cameras' JSONs and imports the capture stack at import time), so the function's own source is compiled out of
    the reference file with ast and RUN here; nothing of it is stored but its outputs."""
    import ast
    tree = ast.parse(open(os.path.join(REF, "CalculateCameraPoses.py")).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "poses_to_fundamental_matrix"][0]
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "CalculateCameraPoses.py", "exec"), ns)
    ref = ns["poses_to_fundamental_matrix"]
    after = json.load(open(os.path.join(REF, "jsons/after_ba_extrinsics.json")))
    cp = json.load(open(os.path.join(REF, "jsons/camera-params-in.json")))
    rec = {}
    poses = [{"R": np.array(p["R"]), "t": np.array(p["t"])} for p in after]
    K = [np.array(c["intrinsic_matrix"]) for c in cp]
    rec["shipped_R"] = np.stack([p["R"] for p in poses]); rec["shipped_t"] = np.stack([p["t"].ravel() for p in poses])
    rec["shipped_K"] = np.stack(K[:2])
    rec["shipped_F"] = ref(poses[0], poses[1], K[0], K[1])
    rec["shipped_E"] = ref(poses[0], poses[1])
    rig = S.config_rig("c3")
    rp = rig["poses"]
    rK = [np.array(c["intrinsic_matrix"], dtype=np.float64) for c in rig["camera_params"]]
    rec["ring_R"] = np.stack([np.asarray(p["R"], dtype=np.float64) for p in rp])
    rec["ring_t"] = np.stack([np.asarray(p["t"], dtype=np.float64).ravel() for p in rp])
    rec["ring_K"] = np.stack(rK)
    rec["ring_F"] = np.stack([ref({"R": rec["ring_R"][0], "t": rec["ring_t"][0]}, {"R": rec["ring_R"][i], "t": rec["ring_t"][i].reshape(3, 1)},
                                  rK[0], rK[i]) for i in range(1, len(rp))])
    np.savez_compressed(os.path.join(OUT, "kat_fundamentals.npz"), **rec)
    print("fundamentals golden:", rec["shipped_F"].shape, rec["ring_F"].shape)


def main():
    import cv2
    os.makedirs(OUT, exist_ok=True)
    os.chdir(REF)
    sys.path.insert(0, REF)
    with quiet():
        import lib.ImageOperations as IO
        import lib.Helpers as Hh
    IO.fast_cuda_blur = lambda image, kernel_size=5: R.blur5_floor(image)
    cp = json.load(open("jsons/camera-params-in.json"))
    K = np.array(cp[0]["intrinsic_matrix"])
    D = np.array(cp[0]["distortion_coef"])
    meta = {"cv2": cv2.__version__, "numpy": np.__version__,
            "scipy": __import__("scipy").__version__, "numba": __import__("numba").__version__,
            "reference": "RashmikaDushan/MocapV2 mounted at /root/reference",
            "blur_sim_equal": bool(blur_sim_check())}
    print("blur simulator check:", meta["blur_sim_equal"])

    def stage_record(img):
        """Reference _find_dot + the cv2 stage outputs its body produces (ImageOperations.py:38-65)."""
        _, pts = IO._find_dot(img.copy())
        und = cv2.undistort(img, K, D)
        grey = IO.image_filter_gpu(und)
        cs, hier = cv2.findContours(grey, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
        table = []
        for k, c in enumerate(cs):
            m = cv2.moments(c)
            area = cv2.contourArea(c)
            per = cv2.arcLength(c, True)
            keep = bool(per and (4 * np.pi * area / (per * per) > 0.5 and area > 500))
            a00o = cv2.contourArea(c, oriented=True) * 2
            table.append([a00o, m["m10"], m["m01"], m["m00"], per, int(hier[0][k][3]), int(keep)])
        n_lab, lab = cv2.connectedComponents(grey, connectivity=8)
        return {"points": pts, "und_crc": zlib.crc32(und.tobytes()), "bin": np.packbits(grey != 0, axis=1),
                "contours": np.array(table, dtype=np.float64).reshape(-1, 7), "n_blobs": int(n_lab - 1),
                "n_fg": int(np.count_nonzero(grey))}

    # ---------------- KAT: the reference's own golden vector -------------------------------------------
    ip = np.array(json.load(open("jsons/image_points.json")))
    obj = np.array(json.load(open("jsons/after_ba_objects.json")))
    with quiet():
        poses, _ = Hh.get_extrinsics()
        Hh.camera_params = None
        tri = Hh.triangulate_points(ip, poses)
        err = Hh.calculate_reprojection_errors(ip, tri, poses)
    np.savez_compressed(os.path.join(OUT, "kat_triangulate.npz"), image_points=ip, objects_json=obj,
                        objects_ref=tri, errors_ref=err,
                        R=np.stack([p["R"] for p in poses]), t=np.stack([p["t"] for p in poses]),
                        K=np.stack([np.array(c["intrinsic_matrix"]) for c in cp]),
                        dist=np.stack([np.array(c["distortion_coef"]) for c in cp]),
                        F=np.array(json.load(open("jsons/fundamentals.json"))))
    print("KAT max |ref - json| =", np.abs(tri - obj).max())
    kat_bundle_adjustment(Hh, ip)

    # ---------------- C1: 2 cams 640x480, 4 markers, shipped calibration, reference as-is --------------
    rig = S.config_rig("c1")
    rng = np.random.default_rng(S.SEED0 + 1000)
    X0 = S.config_markers("c1", rig, rng)
    c1 = {"frames": [], "records": [], "corr": []}
    frames = []
    for f in range(3):
        X = X0 + rng.uniform(-0.01, 0.01, X0.shape)
        radii = rng.integers(14, 23, (2, len(X)))
        fs = S.render_frameset(rig, X, radii, rng)
        frames.append(fs)
        recs = [stage_record(fs[c]) for c in range(2)]
        pts = [copy.deepcopy(r["points"]) for r in recs]
        Hh.camera_params = None
        Hh.Fs = []
        with quiet():
            o, ipa = Hh.find_point_correspondance_and_object_points(copy.deepcopy(pts), poses, 4)
        c1["records"].append(recs)
        c1["corr"].append({"points": pts, "object_points": o.tolist(), "image_points_all": ipa.tolist(), "X": X.tolist()})
    frames = np.stack(frames)
    np.savez_compressed(os.path.join(OUT, "c1_frames.npz"), frames=frames,
                        **{f"bin_{f}_{c}": c1["records"][f][c]["bin"] for f in range(3) for c in range(2)},
                        **{f"contours_{f}_{c}": c1["records"][f][c]["contours"] for f in range(3) for c in range(2)})
    json.dump({"frames_sha": sha(frames),
               "records": [[{k: v for k, v in r.items() if k not in ("bin", "contours")} for r in recs]
                           for recs in c1["records"]],
               "corr": c1["corr"]}, open(os.path.join(OUT, "c1.json"), "w"))
    print("C1 points:", [r["points"] for r in c1["records"][0]])

    # ---------------- C2: video replay (960x540), selected frames --------------------------------------
    vid_frames, vid_recs, vid_ids = [], [], []
    hist = {}
    for cam in range(1, 7):
        cap = cv2.VideoCapture(f"videos/cam{cam}.mp4")
        k = 0
        while True:
            ok, fr = cap.read()
            if not ok:
                break
            g = cv2.cvtColor(fr, cv2.COLOR_BGR2GRAY)
            _, pts = IO._find_dot(g.copy())
            n = 0 if pts == [[None, None]] else len(pts)
            hist[n] = hist.get(n, 0) + 1
            take = (n > 0 and len([i for i in vid_ids if i[0] == cam]) < 6) or (k in (0, 97) and cam in (1, 4, 6))
            if take:
                vid_frames.append(g)
                vid_recs.append(stage_record(g))
                vid_ids.append((cam, k))
            k += 1
    vid_frames = np.stack(vid_frames)
    np.savez_compressed(os.path.join(OUT, "c2_video.npz"), frames=vid_frames, ids=np.array(vid_ids),
                        **{f"bin_{i}": r["bin"] for i, r in enumerate(vid_recs)},
                        **{f"contours_{i}": r["contours"] for i, r in enumerate(vid_recs)})
    json.dump({"hist_all_1200_frames": hist,
               "records": [{k: v for k, v in r.items() if k not in ("bin", "contours")} for r in vid_recs]},
              open(os.path.join(OUT, "c2_video.json"), "w"))
    print("C2 kept-blob histogram over all frames:", hist, "stored", len(vid_ids))

    # ---------------- shapes: rings / holes / edge-touching / nested blobs, odd sizes ---------------------
    rng = np.random.default_rng(S.SEED0 + 77)
    shp_frames, shp_recs = [], []
    for it in range(6):
        H, W = [(131, 97), (200, 333), (480, 640), (64, 64), (300, 301), (257, 511)][it]
        img = rng.integers(0, 40, (H, W)).astype(np.uint8)
        for _ in range(int(rng.integers(2, 7))):
            c = (int(rng.integers(-10, W + 10)), int(rng.integers(-10, H + 10)))
            r = int(rng.integers(12, 40))
            th = -1 if rng.random() < 0.5 else int(rng.integers(6, 14))
            cv2.circle(img, c, r, 255, th)
        img = cv2.GaussianBlur(img, (0, 0), 1.2)
        shp_frames.append(img)
        shp_recs.append(stage_record(img))
    np.savez_compressed(os.path.join(OUT, "shapes.npz"),
                        **{f"frame_{i}": f for i, f in enumerate(shp_frames)},
                        **{f"bin_{i}": r["bin"] for i, r in enumerate(shp_recs)},
                        **{f"contours_{i}": r["contours"] for i, r in enumerate(shp_recs)})
    json.dump({"records": [{k: v for k, v in r.items() if k not in ("bin", "contours")} for r in shp_recs]},
              open(os.path.join(OUT, "shapes.json"), "w"))
    print("shapes points:", [r["points"] for r in shp_recs])

    # ---------------- big frames by seed (inputs regenerated in the test, hash-checked) ------------------
    big = []
    for name, nm in (("c3", 32), ("c4", 128)):
        rig = S.config_rig(name)
        rng = np.random.default_rng(S.SEED0 + {"c3": 3000, "c4": 4000}[name])
        X = S.config_markers(name, rig, rng)
        uv = S.marker_pixels(rig, X)
        for cam in (0, len(rig["poses"]) // 2):
            frng = np.random.default_rng(S.SEED0 + {"c3": 3000, "c4": 4000}[name] + cam * 10)
            radii = frng.integers(14, 23, len(X))
            img = S.render_frame(rig["H"], rig["W"], uv[cam], radii, frng)
            rec = stage_record(img)
            big.append({"config": name, "cam": cam, "sha": sha(img), "points": rec["points"],
                        "und_crc": rec["und_crc"], "bin_sha": sha(rec["bin"]), "n_blobs": rec["n_blobs"],
                        "n_fg": rec["n_fg"], "n_contours": int(len(rec["contours"])),
                        "contours_sha": sha(rec["contours"])})
            print(name, cam, "blobs", rec["n_blobs"], "kept", len(rec["points"]))
    json.dump(big, open(os.path.join(OUT, "big_frames.json"), "w"))

    # ---------------- geometry: multi-camera correspondences from the reference ---------------------------
    geo = []
    for name, nm, nfs, seed in (("c1", 4, 6, 11), ("c3", 32, 2, 12), ("c3", 10, 6, 13), ("c4", 6, 4, 14), ("c5", 12, 3, 15)):
        rig = S.config_rig(name)
        rng = np.random.default_rng(seed)
        Hh.camera_params = np.array(rig["camera_params"])
        Hh.Fs = rig["Fs"]
        X0 = S.config_markers(name, rig, rng)[:nm]
        for f in range(nfs):
            X = X0 + rng.uniform(-0.01, 0.01, X0.shape)
            uv = S.marker_pixels(rig, X)
            pts = []
            for c in range(len(rig["poses"])):
                p = [[int(u), int(v)] for u, v in uv[c] + rng.uniform(-1.5, 1.5, (len(X), 2))]
                order = rng.permutation(len(p))
                p = [p[k] for k in order if rng.random() > 0.05]
                if f == nfs - 1 and c == 1 and name != "c1":
                    p = []                               # a camera that saw nothing
                if len(p) == 0:
                    p = [[None, None]]
                pts.append(p)
            with quiet():
                o, ipa = Hh.find_point_correspondance_and_object_points(copy.deepcopy(pts), rig["poses"], nm)
                # per-group reference numbers for the first complete group of every root
            geo.append({"config": name, "obj_count": nm, "points": pts, "object_points": o.tolist(),
                        "image_points_all": ipa.tolist()})
    json.dump(geo, open(os.path.join(OUT, "geometry.json"), "w"))
    print("geometry cases:", len(geo))

    # triangulation / reprojection on many-view groups from the reference (C5-like 8 views)
    rig = S.config_rig("c5")
    rng = np.random.default_rng(99)
    Hh.camera_params = np.array(rig["camera_params"])
    X = S.config_markers("c5", rig, rng)[:64]
    uv = S.marker_pixels(rig, X) + rng.uniform(-2, 2, (8, len(X), 2))
    groups = np.transpose(np.floor(uv), (1, 0, 2))
    with quiet():
        tri = Hh.triangulate_points(groups, rig["poses"])
        err = Hh.calculate_reprojection_errors(groups, tri, rig["poses"])
    np.savez_compressed(os.path.join(OUT, "c5_groups.npz"), groups=groups, objects_ref=tri, errors_ref=err)
    json.dump(meta, open(os.path.join(OUT, "meta.json"), "w"), indent=1)
    print("done ->", OUT)


if __name__ == "__main__":
    if sys.argv[1:] == ["fundamentals"]:
        kat_fundamentals()
    else:
        main()
        kat_fundamentals()
