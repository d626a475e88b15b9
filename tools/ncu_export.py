"""Turns a raw-page export of an `ncu --set full` capture (`ncu -i X.ncu-rep --page raw --csv > X_raw.csv`) into the
tracked evidence under profiles/:

    python tools/ncu_export.py gpurun_out/r2c_ncu_dom_raw.csv profiles/r2_ncu_dominant --dominant conv_tc_halo_kernel --batch 32

writes  <out>_raw.csv   the raw page itself (one row per captured launch, every metric of the full set),
        <out>.md        a readable per-launch table of the metrics the roofline argument uses,
and with --dominant K also profiles/ncu_dominant_kernel.json, which bench.py reads for `roofline.traffic`
(dram__bytes_read.sum + dram__bytes_write.sum of the first captured launch of kernel K at `--batch` samples)."""
import argparse
import csv
import json
import os
import shutil

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "DRAM rd"),
    ("dram__bytes_write.sum", "DRAM wr"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp inst"),
]


def to_bytes(v, unit):
    m = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * m.get(unit, 1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw_csv")
    ap.add_argument("out_prefix")
    ap.add_argument("--dominant", default=None)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--title", default="")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.raw_csv)))
    hdr, units, data = rows[0], rows[1], [r for r in rows[2:] if len(r) == len(rows[0])]
    ix = {h: i for i, h in enumerate(hdr)}
    os.makedirs(os.path.dirname(a.out_prefix) or ".", exist_ok=True)
    shutil.copyfile(a.raw_csv, a.out_prefix + "_raw.csv")
    cols = [(c, n) for c, n in COLS if c in ix]
    with open(a.out_prefix + ".md", "w") as fh:
        fh.write(f"# {a.title or os.path.basename(a.out_prefix)}\n\nSource: `{os.path.basename(a.out_prefix)}_raw.csv` "
                 f"(raw page of an `ncu --set full --clock-control none` capture, one row per launch).\n\n")
        fh.write("| # | kernel | grid | " + " | ".join(n for _, n in cols) + " |\n")
        fh.write("|---|---|---|" + "---|" * len(cols) + "\n")
        for i, r in enumerate(data):
            name = r[ix["Kernel Name"]].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
            cells = []
            for c, _ in cols:
                v, u = r[ix[c]], units[ix[c]]
                try:
                    cells.append(f"{float(v):.4g} {u}".strip())
                except ValueError:
                    cells.append(v)
            fh.write(f"| {i} | `{name}` | {r[ix['Grid Size']]} | " + " | ".join(cells) + " |\n")
    if a.dominant:
        for r in data:
            if a.dominant in r[ix["Kernel Name"]]:
                rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
                wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
                out = {"kernel": a.dominant, "batch": a.batch, "dram_bytes_read": rd, "dram_bytes_write": wr,
                       "duration_ms_under_ncu": float(r[ix["gpu__time_duration.sum"]]) *
                       {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[ix["gpu__time_duration.sum"]], 1.0),
                       "source": os.path.relpath(a.out_prefix + "_raw.csv")}
                path = os.path.join(os.path.dirname(a.out_prefix) or ".", "ncu_dominant_kernel.json")
                json.dump(out, open(path, "w"), indent=1)
                print("wrote", path, out)
                break
    print("wrote", a.out_prefix + ".md")


if __name__ == "__main__":
    main()
