"""`ncu -i X.ncu-rep --page raw --csv` -> (1) a per-kernel table (duration, DRAM bytes, registers, occupancy, issue rate, FP64
pipe, top stall reasons) and (2) the DRAM-traffic-per-launch file bench.py reads (profiles/ncu_traffic.json).
usage: ncu_summary.py raw.csv kernels_out.json [traffic_out.json "source text"]"""
import csv
import json
import sys

FAM = [("k_obs_residual", "obs_residual"), ("k_select", "select_median"), ("k_obs_assemble", "obs_assemble"),
       ("k_dynamics_stm", "dynamics_stm"), ("k_quat_terms", "quat_terms"), ("k_system", "system_build"),
       ("k_chain_forward", "blocktridiag_solve"), ("k_chain_backward", "blocktridiag_backsub"),
       ("k_solve_init", "solve_init"), ("k_retract", "retract"), ("k_obs_trial", "trial_residual(obs)"),
       ("k_dyn_trial", "trial_residual(dyn)"), ("k_accept", "accept_reduce(accept)"), ("k_init_residual", "accept_reduce(init)"),
       ("k_resjac", "project_resjac"), ("k_satcam_visibility", "satcam_visibility")]


def fam(name):
    n = name.replace("void ", "").replace("vs::", "")
    for k, f in FAM:
        if n.startswith(k):
            return f
    return n.split("(")[0]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = rows[0]
    out, traffic = {}, {}
    stall = [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
    g = lambda d, k: float(d[k]) if d.get(k) not in (None, "", "n/a") else None
    units = dict(zip(hdr, rows[1]))

    def to_bytes(d, k):
        v, u = g(d, k), units.get(k, "")
        return None if v is None else v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)

    def to_ms(d, k):
        v, u = g(d, k), units.get(k, "")
        return None if v is None else v * {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        f = fam(d["Kernel Name"])
        if f in out:
            continue
        rd, wr = to_bytes(d, "dram__bytes_read.sum"), to_bytes(d, "dram__bytes_write.sum")
        st = sorted(((g(d, h) or 0.0, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for h in stall), reverse=True)[:5]
        out[f] = {"kernel": d["Kernel Name"].replace("void ", "").split("(")[0], "duration_ms": to_ms(d, "gpu__time_duration.sum"),
                  "dram_bytes_read": rd, "dram_bytes_write": wr, "registers": g(d, "launch__registers_per_thread"),
                  "grid": d.get("launch__grid_size"), "block": d.get("launch__block_size"),
                  "warps_active_pct": g(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                  "issue_active_per_cycle": g(d, "smsp__issue_active.avg.per_cycle_active"),
                  "fp64_pipe_pct": g(d, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                  "warp_instructions": g(d, "smsp__inst_executed.sum"), "top_stalls": {n: round(v, 2) for v, n in st}}
        if rd is not None and wr is not None:
            traffic[f] = {"dram_bytes_per_launch": int(rd + wr), "kernel": out[f]["kernel"]}
    json.dump({"capture": "ncu --set full --clock-control none --import-source on, one launch per kernel", "kernels": out},
              open(sys.argv[2], "w"), indent=1)
    if len(sys.argv) > 4:
        # bench.py looks families up by its own timing-family names
        alias = {"trial_residual(obs)": "trial_residual", "accept_reduce(accept)": "accept_reduce"}
        for k, v in list(traffic.items()):
            if k in alias and alias[k] not in traffic:
                traffic[alias[k]] = v
        json.dump({"source": sys.argv[4], "families": traffic}, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
