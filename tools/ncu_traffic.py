"""dram bytes per launch per kernel family from an ncu summary CSV (tools/ncu_summary.py) -> JSON fragment for profiles/r02_roofline_traffic.json:
   python tools/ncu_traffic.py profiles/r02_ncu_full_ragged4096.csv ragged4096"""
import csv
import json
import sys

FAMILY = [("EpiQKV", "gemm_qkv"), ("attention", "attention"), ("gemm_ln_kernel<0", "gemm_out_proj"), ("gemm_ln_kernel<2", "gemm_out_proj"),
          ("gemm_ln_kernel<1", "gemm_ffn2"), ("EpiOperand", "gemm_ffn1"), ("fbank_kernel", "fbank"), ("ctc_greedy", "ctc_greedy"), ("beam_kernel", "beam"),
          ("EpiF32", "gemm_f32")]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
i_n, i_r, i_w, i_t = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
acc = {}
for r in rows[2:]:
    fam = next((f for k, f in FAMILY if k in r[i_n]), None)
    if fam is None:
        continue
    b = float(r[i_r]) * UNIT[units[i_r]] + float(r[i_w]) * UNIT[units[i_w]]
    acc.setdefault(fam, []).append(b)
print(json.dumps({sys.argv[2]: {k: int(sum(v) / len(v)) for k, v in acc.items()}}, indent=1))
