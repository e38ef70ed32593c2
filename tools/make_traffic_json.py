"""profiles/traffic.json (read by bench.py for `roofline.traffic`) from the ncu --set full summaries under profiles/."""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")
SOURCES = {
    # steady state: bench.py itself under ncu with --cache-control none, launches 41..43 of the rotating-buffer loop, so the write-backs
    # of earlier launches' tau are part of what each launch moves
    "rnea_f64_1048576": "r2_rnea_f64_steady_ncu_full.csv",
    "rnea_f32_1048576": "r1c_rnea_f32_ncu_full.csv",
    "gram_f64_12500000": "r2_gram_v3_ncu_full.csv",
    "gram_f32_12500000": "r2_gram32_v3_ncu_full.csv",
    "linearize_f64_1048576": "r2_lin_final_ncu_full.csv",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    out, notes = {}, {}
    for key, fname in SOURCES.items():
        rows = {r[0]: r[1:] for r in csv.reader(open(os.path.join(PROF, fname)))}
        rd, wr = rows["dram__bytes_read.sum"], rows["dram__bytes_write.sum"]
        n = len(rd) - 1
        tot = [float(rd[1 + i]) * UNIT[rd[0]] + float(wr[1 + i]) * UNIT[wr[0]] for i in range(n)]
        out[key] = sum(tot) / n
        notes[key] = f"dram__bytes_read.sum + dram__bytes_write.sum, mean of {n} launch(es), profiles/{fname}"
    out["_source"] = notes
    out["_note"] = ("bytes per launch from ncu --set full (dram__bytes_read.sum + dram__bytes_write.sum).  rnea_f64 is a steady-state capture "
                    "(no cache flush between the rotating-buffer launches): 124 MB read + 50 MB written = the 168 B/sample the kernel really "
                    "moves (15 live input rows + 6 output rows), against 192 B/sample of SURVEY 8(d)'s contract bytes")
    with open(os.path.join(PROF, "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print({k: v for k, v in out.items() if not k.startswith("_")})


if __name__ == "__main__":
    main()
