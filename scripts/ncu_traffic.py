"""Per-family DRAM traffic of ONE train_batch from an ncu launch list.

  gpurun -- 'python bench.py --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline > /dev/null &&
             ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
                 -k regex:conv_tc --csv --page raw --log-file gpurun_out/r2_traffic.csv \
                 python bench.py --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline > /dev/null'
  python scripts/ncu_traffic.py gpurun_out/r2_traffic.csv profiles/conv_tc_traffic.json

bench.py runs 1 warm-up pair of host-launched steps, one probe step, `warmup` + `steps` + 1 timed steps, so the capture
holds a whole number of identical steps; the script divides by the number of steps it finds (launches of the family per
step is printed by bench.py as roofline.launches).  The output holds the MEAN dram bytes per launch of the family
(fwd + dgrad kernels: conv_tc_fwd_kernel, conv_tc_halo_kernel, conv_tc_halo2_kernel) -- directly comparable with
roofline.algorithmic_bytes_per_launch_step_mean, which is the mean over the same launches.
"""
import csv
import json
import re
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    fam = re.compile(sys.argv[3] if len(sys.argv) > 3 else r"conv_tc_(fwd|halo|halo2)_kernel")
    with open(src, newline="") as f:
        table = [r for r in csv.reader(l for l in f if l.startswith('"'))]
    header, units, data = table[0], dict(zip(table[0], table[1])), table[2:]
    rows = [dict(zip(header, r)) for r in data if len(r) == len(header)]
    sel = [r for r in rows if fam.search(r["Kernel Name"])]
    if not sel:
        raise SystemExit("no launches of the family in %s" % src)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def total(key):
        return sum(float(r[key].replace(",", "")) for r in sel if r[key] != "") * scale[units[key]]
    rd_b, wr_b = total("dram__bytes_read.sum"), total("dram__bytes_write.sum")
    out = {
        "kernel": "conv_tc fwd+dgrad family (conv_tc_fwd_kernel, conv_tc_halo_kernel, conv_tc_halo2_kernel)",
        "dram_bytes_per_launch": (rd_b + wr_b) / len(sel),
        "dram_read_bytes_per_launch": rd_b / len(sel),
        "dram_write_bytes_per_launch": wr_b / len(sel),
        "launches_captured": len(sel),
        "source": "%s (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over EVERY launch of the family in the "
                  "captured host-launched steps; mean per launch, comparable with algorithmic_bytes_per_launch_step_mean)" % src,
    }
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
