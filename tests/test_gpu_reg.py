"""(f)4 on the device: prior_gpu, propagate_dynamics_cov_init and BA_reg against outputs of the UNMODIFIED reference
(tests/golden/ba_reg.npz, tests/golden/make_golden_reg.py) and the oracle pinned to them (oracle/reg_oracle.py)."""
import numpy as np
import pytest
import torch

import reg_oracle as ro
from conftest import load_golden
from vinsat_b200 import _lib
from vinsat_b200.BA import BA_filtering as F
from vinsat_b200.BA import BA_utils as U

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(np.asarray(a) - b).max() / max(np.abs(b).max(), 1e-300)


def test_prior_vs_reference():
    g = load_golden("ba_reg")
    t = torch.tensor
    N = g["pri_states"].shape[0]
    r, Jp, Hqp, qg = U.prior_gpu(t(g["pri_states"])[None], t(g["pri_prop"])[None], 1, 1, t(g["pri_Hs"])[None], t(g["pri_Hr"])[None])
    assert r.shape == (1, N, 7) and Jp.shape == (1, 6 * N, 9 * N) and Hqp.shape == (1, 9 * N, 9 * N) and qg.shape == (1, N, 9)
    assert rel(r[0].numpy(), g["pri_r"]) < 1e-12
    J4 = Jp[0].numpy().reshape(N, 6, N, 9)
    H4 = Hqp[0].numpy().reshape(N, 9, N, 9)
    assert rel(np.stack([J4[i][:, i] for i in range(N)]), g["pri_Jp"]) < 1e-12
    assert sum(np.abs(J4[i][:, j]).max() for i in range(N) for j in range(N) if i != j) == 0
    scale = np.abs(np.einsum("nki,nkj->nij", g["pri_Jp"], g["pri_Jp"])).max()
    assert np.abs(np.stack([H4[i][:, i] for i in range(N)]) - g["pri_Hqp"]).max() < 1e-12 * scale
    assert np.abs(qg[0].numpy() - g["pri_qgrad"]).max() < 1e-12 * np.abs(g["pri_r"]).max()
    r2 = U.prior_gpu(t(g["pri_states"])[None], t(g["pri_prop"])[None], 1, 100, t(g["pri_Hs"])[None], t(g["pri_Hr"])[None],
                     jacobian=False)
    assert rel(r2[0].numpy(), g["pri_r_trial"]) < 1e-12
    z = U.prior_gpu(t(g["pri_states"])[None], t(g["pri_prop"])[None], 1, 1, t(g["pri_Hs"])[None], t(g["pri_Hr"])[None], initialize=True)
    assert z[0].shape == (1, N, 6) and float(z[1].abs().sum()) == 0


def test_covariance_propagation_vs_reference():
    g = load_golden("ba_reg")
    t = torch.tensor
    s_t, v_t, hs, hr = U.propagate_dynamics_cov_init(t(g["cov_state"])[None], t(g["cov_vel"])[None], t(g["cov_hessian"])[None],
                                                     t(g["cov_omega"])[None], int(g["cov_tdiff"]), int(g["cov_duration"]), 1)
    assert rel(s_t[0].numpy(), g["cov_states_t"]) < 1e-12 and rel(v_t[0].numpy(), g["cov_vel_t"]) < 1e-12
    assert rel(hs[0].numpy(), g["cov_hess_state_t"]) < 1e-9 and rel(hr[0].numpy(), g["cov_hess_rot_t"]) < 1e-9


def test_ba_reg_tracks_reference_history(capsys):
    """Six BA_reg iterations (several with LM rejections up to the lamda cap): states within 1 m / 1 mm/s of the
    reference at every iteration, identical lamda schedule, last Hessian 1e-6."""
    g = load_golden("ba_reg")
    pr = {k[7:]: v for k, v in g.items() if k.startswith("reg_in_")}
    t = torch.tensor
    T = pr["states0"].shape[0]
    N = int(np.diff(pr["time_idx"]).max())
    imu = torch.zeros(1, T, N, 10, dtype=torch.float64)
    imu[0, :, -1, 6:10] = t(pr["cum_rot"])
    st, lam = t(g["reg_start"])[None], 1e-4
    vel = t(pr["velocities"])[None]
    st_or, lam_or = g["reg_start"].copy(), 1e-4
    for j in range(len(g["reg_lamda_hist"])):
        it = int(g["reg_first_iter"]) + j
        st, _, lam, H = F.BA_reg(it, st, vel, t(g["reg_prior"])[None], vel, t(g["reg_Hs"])[None], t(g["reg_Hr"])[None], imu,
                                 t(pr["uv"])[None], t(pr["xyz"])[None], pr["ii"], pr["time_idx"], t(pr["intr"])[None],
                                 t(pr["conf"]), 1e-3, 1e-3, lam, t(pr["states_gt"][:, :7]))
        st_or, lam_or, H_or, info = ro.ba_reg_iteration(it, st_or, g["reg_prior"], g["reg_Hs"], g["reg_Hr"], pr["cum_rot"],
                                                        pr["uv"], pr["xyz"], pr["ii"], pr["time_idx"], pr["intr"], pr["conf"],
                                                        lam_or)
        s, ref = st[0].numpy(), g["reg_states_hist"][j]
        assert np.abs(s[:, :3] - ref[:, :3]).max() < 1e-3, j
        assert np.abs(s[:, 7:] - ref[:, 7:]).max() < 1e-6, j
        assert np.abs(s[:, 3:7] - ref[:, 3:7]).max() < 1e-7, j
        assert lam == g["reg_lamda_hist"][j], j
        assert rel(H[0].numpy(), g["reg_hessian_hist"][j]) < 1e-6, j
    # a plain BA() call on the same cached batch afterwards is unaffected by the prior
    s2, _, _, _ = F.BA(18, st, vel, imu, t(pr["uv"])[None], t(pr["xyz"])[None], pr["ii"], pr["time_idx"], t(pr["intr"])[None],
                       t(pr["conf"]), 1e-3, 1e-3, 1e-4, t(pr["states_gt"][:, :7]))
    import ba_oracle as o
    so_, _, _, _ = o.ba_iteration(18, st[0].numpy(), pr["cum_rot"], pr["uv"], pr["xyz"], pr["ii"], pr["time_idx"], pr["intr"],
                                  pr["conf"], 1e-4)
    assert np.abs(s2[0].numpy()[:, :3] - so_[:, :3]).max() < 1e-3
