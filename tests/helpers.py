"""Shared helpers of the test-suite (oracle-side utilities; never imported by the product package)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_kats():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


def inf_list(x):
    return np.array([[np.inf if v == "inf" else v for v in row] for row in x], dtype=np.float64)


def kat_iterator(o, kat):
    if kat["iterator"] == "product":
        return o.product_iterator(kat["nu"])
    _, lb, ub = kat["iterator"]
    return o.bounded_sum_iterator(kat["nu"], lb, ub)


def load_seeded():
    z = np.load(os.path.join(GOLDEN, "seeded.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    out = {}
    for nm in names:
        meta = json.loads(str(z[nm + "/meta"]))
        meta["p"] = float("inf") if meta["p"] == "inf" else meta["p"]
        meta["iterator"] = [tuple(t) for t in meta["iterator"]]
        out[nm] = dict(meta=meta, **{k.split("/")[1]: z[k] for k in z.files if k.startswith(nm + "/") and not k.endswith("meta")})
    return out


def phi_admissible(o, Phi, nu, iterator):
    """(2, K, B+1) view of a reference-shaped Phi restricted to the admissible levels."""
    g = o.grid_offsets(nu, iterator)
    B1 = Phi.shape[-1]
    return Phi.reshape(2, -1, B1)[:, g, :]


def random_instance(rng, o, *, tie_heavy=False, K_choice=None, n_max=8, B_max=9):
    """Small random instance in the style of SURVEY 8c properties."""
    kind = K_choice if K_choice is not None else rng.choice([3, 5, 6, 36])
    if kind == 3:
        nu = [[0, 1]] * 3; it = o.bounded_sum_iterator(nu, 1, 1)
    elif kind == 5:
        nu = [[-2, -1, 0, 1, 2]]; it = o.product_iterator(nu)
    elif kind == 6:
        nu = [[0, 1, 2], [0, 1]]; it = o.product_iterator(nu)
    else:
        nu = [[0, 1, 2, 3, 4, 5]] * 2; it = o.product_iterator(nu)
    n = int(rng.integers(1, n_max + 1))
    B = int(rng.integers(0, B_max + 1))
    M = len(nu)
    df = rng.standard_normal((n, M))
    beta = float(rng.uniform(0.01, 1.0))
    if tie_heavy:
        df = np.round(df * 4) / 4
        beta = 0.25
    lv = o.level_values(nu, it)
    u_old = lv[rng.integers(0, len(it), size=n)].astype(np.float64)
    p = [1, 2, float("inf")][int(rng.integers(0, 3))]
    dt = float(rng.choice([1.0, 0.5, 0.125]))
    return dict(nu=[list(v) for v in nu], it=it, n=n, B=B, df=df, u_old=u_old, beta=beta, p=p, dt=dt)


def objective_of(o, u, inst, cost):
    """Subproblem objective of a trajectory, summed in trajectory order (tolerance compare only)."""
    lv = o.level_values(inst["nu"], inst["it"])
    ks = [int(np.where((lv == u[i]).all(axis=1))[0][0]) for i in range(u.shape[0])]
    val = sum(float(inst["dt"] * inst["df"][i] @ lv[ks[i]]) for i in range(u.shape[0]))
    val += sum(cost[ks[i + 1], ks[i]] for i in range(u.shape[0] - 1))
    return val
