"""ncu launch list (gpu__time_duration.sum per launch) of `bench.py --steps K --warmup W` -> per-family share of a
timed step, next to the CUDA-event shares of the plain run.  usage: summarize_launches.py launches.csv bench.json W K"""
import csv, json, sys

FAM = [("k_obs_residual", "obs_residual"), ("k_select", "select_median"), ("k_obs_assemble", "obs_assemble"),
       ("k_dynamics_stm", "dynamics_stm"), ("k_quat_terms", "quat_terms"), ("k_system", "system_build"),
       ("k_chain_forward", "blocktridiag_solve"), ("k_chain_backward", "blocktridiag_backsub"),
       ("k_solve_init", "solve_init"), ("k_retract", "retract"), ("k_obs_trial", "trial_residual"),
       ("k_dyn_trial", "trial_residual"), ("k_accept", "accept_reduce"), ("k_init_residual", "accept_reduce")]


def family(name):
    for key, fam in FAM:
        if name.startswith(key):
            return fam
    return "other:" + name


def main():
    path, bench, W, K = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    names = [r[4].replace("void ", "").replace("vs::", "").split("(")[0] for r in rows]
    ns = [float(r[-1]) for r in rows]
    starts = [i for i, n in enumerate(names) if n.startswith("k_obs_residual")]
    lo = starts[W]
    hi = starts[W + K] if len(starts) > W + K else len(names)
    while hi > lo and not names[hi - 1].startswith("k_accept"):      # drop the upload kernels of the next phase
        hi -= 1
    tot = {}
    cnt = {}
    for n, t in zip(names[lo:hi], ns[lo:hi]):
        f = family(n)
        tot[f] = tot.get(f, 0.0) + t * 1e-6 / K
        cnt[f] = cnt.get(f, 0) + 1
    T = sum(tot.values())
    ev = json.load(open(bench))["extra"]["kernel_ms_per_step"]
    TE = sum(ev.values())
    out = {"launches_per_step": (hi - lo) / K, "ncu_ms_per_step": round(T, 3), "event_ms_per_step": round(TE, 3), "families": {}}
    for f in sorted(tot, key=lambda k: -tot[k]):
        out["families"][f] = {"ncu_ms": round(tot[f], 3), "ncu_share": round(tot[f] / T, 4), "launches_per_step": cnt[f] / K,
                              "event_ms": ev.get(f), "event_share": round(ev.get(f, 0.0) / TE, 4)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
