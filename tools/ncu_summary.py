"""Extracts the metrics DESIGN.md / bench.py cite from an ncu report (`ncu -i X.ncu-rep --page raw --csv`) into a small JSON.

    python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/X_summary.json [launch index]
"""
import csv
import io
import json
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2 + which]
    d = {h: {"unit": u, "value": v} for h, u, v in zip(hdr, units, vals)}
    res = {"_kernel": d.get("Kernel Name", {}).get("value"), "_report": rep}
    for k in WANT:
        if k in d:
            res[k] = d[k]
    for k in sorted(d):
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            res[k] = d[k]
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: v["value"] for k, v in res.items() if isinstance(v, dict)}, indent=0)[:1500])


if __name__ == "__main__":
    main()
