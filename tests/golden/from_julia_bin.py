"""Reader of the `.bin` files julia/dump_golden.jl writes from the REAL reference (HybridVariationalInference.jl's own
neg_elbo_gtf_components + Zygote gradient with replayed draws).

    python tests/golden/from_julia_bin.py tests/golden/julia/doublemm_default.bin      # prints a summary, writes .npz

Layout (flat little-endian float64): header of 10 values [magic 20260, n_phi, n_covar, n_obs, n_batch, n_MC, n_P, n_M,
sizeof(eltype ϕ), n_site_total], then ϕ | xM | xP | y_o | y_unc | i_sites (one-based) | zP | zMs | nL | 6 components | ∇ϕ,
matrices column-major.  `write_bin` is the inverse (used by the round-trip test, which needs no Julia)."""
import sys

import numpy as np

MAGIC = 20260.0
FIELDS = ("phi", "xM", "xP", "y_o", "y_unc", "i_sites", "zP", "zMs", "nL", "comps", "grad")


def _shapes(h):
    n_phi, nC, nO, nb, M, nP, nM = (int(h[k]) for k in ("n_phi", "n_covar", "n_obs", "n_batch", "n_MC", "n_P", "n_M"))
    return dict(phi=(n_phi,), xM=(nC, nb), xP=(2 * nO, nb), y_o=(nO, nb), y_unc=(nO, nb), i_sites=(nb,), zP=(M, nP),
                zMs=(M, nM * nb), nL=(1,), comps=(6,), grad=(n_phi,))


def load_bin(path):
    raw = np.fromfile(path, dtype="<f8")
    if raw.size < 10 or raw[0] != MAGIC:
        raise ValueError(f"{path}: not a dump_golden.jl file (magic {raw[0] if raw.size else None})")
    names = ("magic", "n_phi", "n_covar", "n_obs", "n_batch", "n_MC", "n_P", "n_M", "elsize", "n_site_total")
    h = dict(zip(names, raw[:10]))
    out, pos = {"header": h, "dtype": "f32" if int(h["elsize"]) == 4 else "f64"}, 10
    for k, shp in _shapes(h).items():
        n = int(np.prod(shp))
        if pos + n > raw.size:
            raise ValueError(f"{path}: truncated at field {k}")
        out[k] = np.asfortranarray(raw[pos:pos + n].reshape(shp, order="F"))
        pos += n
    if pos != raw.size:
        raise ValueError(f"{path}: {raw.size - pos} trailing values")
    out["nL"] = float(out["nL"][0])
    out["i_sites0"] = out["i_sites"].astype(np.int64) - 1          # zero-based for the C ABI
    return out


def write_bin(path, G, elsize=8, n_site_total=None):
    nC, nb = G["xM"].shape
    M, nP = G["zP"].shape
    hdr = [MAGIC, G["phi"].size, nC, G["y_o"].shape[0], nb, M, nP, G["zMs"].shape[1] // nb, elsize,
           n_site_total if n_site_total is not None else nb]
    parts = [np.asarray(hdr, dtype="<f8")]
    i_sites = G.get("i_sites", np.arange(1, nb + 1))
    for k in FIELDS:
        a = i_sites if k == "i_sites" else G[k]
        parts.append(np.asarray(a, dtype="<f8").ravel(order="F"))
    np.concatenate(parts).tofile(path)


if __name__ == "__main__":
    for p in sys.argv[1:]:
        G = load_bin(p)
        print(p, G["dtype"], {k: int(v) for k, v in G["header"].items()}, "nL =", G["nL"])
        np.savez_compressed(p[:-4] + ".npz", **{k: G[k] for k in FIELDS})
