"""`ncu --page raw --csv` export(s) -> profiles/ncu_traffic.json: measured DRAM bytes (read + write) per launch of the hot-path
kernels, keyed "<kernel>[<S1|S2>]".  bench.py reports the figure as roofline.traffic next to the algorithmic bytes.
usage: python tools/ncu_traffic.py S1=gpurun_out/prof_s1_raw.csv S2=gpurun_out/prof_s2_raw.csv"""
import csv
import json
import os
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
NAMES = {"vil_tc_fwd5_kernel": "fwd", "vil_tc_fwd3_kernel": "fwd", "addnorm_fwd": "addnorm_fwd", "addnorm_bwd": "addnorm_bwd",
         "bias_act_fwd": "bias_gelu_fwd", "bias_act_bwd<__nv_bfloat16, 1>": "bias_gelu_bwd", "bias_act_bwd<__nv_bfloat16, 0>": "colsum", "vil_tc_fwd2_kernel": "fwd_variant2", "vil_tc_bwd2_dq_kernel": "bwd_dq", "vil_tc_bwd2_dkv_kernel": "bwd_dkv",
         "vil_tc_fwd2_merge": "fwd_merge", "vil_tc_bwd2_merge": "bwd_merge",
         "vil_tc_fwd_kernel": "round1_fwd_local", "vil_tc_bwd_dq_kernel": "round1_bwd_dq", "vil_tc_bwd_dkv_kernel": "round1_bwd_dkv"}
out = {}
for arg in sys.argv[1:]:
    tag, path = arg.split("=")
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        for k, short in NAMES.items():
            if k in name:
                rd = float(r[ir].replace(",", "")) * UNIT[units[ir]]
                wr = float(r[iw].replace(",", "")) * UNIT[units[iw]]
                if f"{short}[{tag}]" in out:
                    continue
                out[f"{short}[{tag}]"] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr,
                                          "ncu_duration": r[it] + " " + units[it], "kernel": name.split("(")[0]}
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
