"""Derive profiles/traffic.json and profiles/fp64_counts.json from ncu reports (read here, no GPU needed):

    python tools/profile_tables.py c2=gpurun_out/r02_prof_c2.ncu-rep c3=... [c5x8=gpurun_out/r02_prof_c5.ncu-rep]

For every workload: DRAM bytes (read + write) and FP64 flops (2 DFMA + DMUL + DADD, from the sass op counters) of the
profiled launches, summed over the kernels of one assembly.  `c5x8=` scales a 128^3 hex capture to the 256^3 workload
(8x the elements; ncu's save / restore of 44 GB per replay pass is avoided)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def num(x):
    return float(x.replace(",", ""))


def main():
    traffic, flops, notes = {}, {}, {}
    for spec in sys.argv[1:]:
        key, rep = spec.split("=")
        scale = 1.0
        if key.endswith("x8"):
            key, scale = key[:-2], 8.0
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        u = dict(zip(hdr, units))
        tb = fl = 0.0
        names = []
        for data in rows[2:]:
            rec = dict(zip(hdr, data))
            names.append(rec["Kernel Name"].split("(")[0])
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tb += num(rec[k]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u[k]]
            f = lambda op: num(rec[f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed"])
            fl += (2.0 * f("dfma") + f("dmul") + f("dadd")) * num(rec["smsp__cycles_elapsed.avg"])
        name = f"{key}:gather"
        traffic[name] = int(tb * scale)
        flops[name] = int(fl * scale)
        notes[name] = f"{os.path.basename(rep)}: {' + '.join(names)}" + (f", scaled x{scale:g} to the named size" if scale != 1 else "")
    for fname, table, what in (("traffic.json", traffic, "dram__bytes_read.sum + dram__bytes_write.sum"),
                               ("fp64_counts.json", flops, "2 * DFMA + DMUL + DADD thread instructions (smsp__sass_thread_inst_executed_op_*_pred_on)")):
        path = os.path.join(ROOT, "profiles", fname)
        table = dict(table)
        table["_source"] = f"ncu --set full, one assembly per workload: {what}; " + "; ".join(f"{k}: {v}" for k, v in notes.items())
        with open(path, "w") as f:
            json.dump(table, f, indent=1)
            f.write("\n")
        print(path, {k: v for k, v in table.items() if not k.startswith("_")})


if __name__ == "__main__":
    main()
