"""Parity metrics between two renders of the same scene (SURVEY.md §8d "Parity tolerances")."""
import numpy as np

ID_AGREE_MIN = 0.999        # hit ids bit-exact on >= 99.9 % of the pixels
DEPTH_TOL = 1e-4            # |dd| <= 1e-4 * max(1, t) on id-agreeing pixels
NORMAL_TOL = 1e-5           # max-abs on id-agreeing pixels
PSNR_MIN = 50.0             # colour, float frame clamped to [0,1]


def compare(a, b):
    """a, b: dicts with depth (n,), normal (n,3), color (n,3), hit_id (n,). Returns a metrics dict."""
    n = a["hit_id"].size
    ida, idb = a["hit_id"].reshape(-1), b["hit_id"].reshape(-1)
    agree = ida == idb
    m = {"pixels": int(n), "id_agree": float(agree.mean()), "id_mismatch": int((~agree).sum())}
    da, db = a["depth"].reshape(-1)[agree], b["depth"].reshape(-1)[agree]
    fin = np.isfinite(da) & np.isfinite(db)
    m["sentinel_mismatch"] = int((np.isfinite(da) != np.isfinite(db)).sum())
    if fin.any():
        err = np.abs(da[fin].astype(np.float64) - db[fin]) / np.maximum(1.0, np.abs(db[fin]))
        m["depth_max_rel"] = float(err.max())
    else:
        m["depth_max_rel"] = 0.0
    na, nb = a["normal"].reshape(-1, 3)[agree], b["normal"].reshape(-1, 3)[agree]
    m["normal_max_abs"] = float(np.abs(na.astype(np.float64) - nb).max()) if len(na) else 0.0
    ca = np.clip(np.nan_to_num(a["color"].reshape(-1, 3).astype(np.float64)), 0, 1)
    cb = np.clip(np.nan_to_num(b["color"].reshape(-1, 3).astype(np.float64)), 0, 1)
    mse = float(((ca - cb) ** 2).mean())
    m["color_max_abs"] = float(np.abs(ca - cb).max())
    m["color_max_abs_id_agree"] = float(np.abs(ca[agree] - cb[agree]).max()) if agree.any() else 0.0
    m["color_psnr"] = float(10 * np.log10(1.0 / mse)) if mse > 0 else float("inf")
    return m


def assert_parity(m, what=""):
    assert m["id_agree"] >= ID_AGREE_MIN, f"{what}: hit ids agree on {m['id_agree']:.5f} < {ID_AGREE_MIN} ({m})"
    assert m["sentinel_mismatch"] == 0, f"{what}: miss sentinels differ ({m})"
    assert m["depth_max_rel"] <= DEPTH_TOL, f"{what}: depth error {m['depth_max_rel']:.3g} > {DEPTH_TOL} ({m})"
    assert m["normal_max_abs"] <= NORMAL_TOL, f"{what}: normal error {m['normal_max_abs']:.3g} > {NORMAL_TOL} ({m})"
    assert m["color_psnr"] >= PSNR_MIN, f"{what}: colour PSNR {m['color_psnr']:.2f} dB < {PSNR_MIN} ({m})"
