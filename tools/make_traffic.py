#!/usr/bin/env python
"""Adds the DRAM traffic of the first profiled launch of an .ncu-rep to profiles/traffic.json (read by bench.py for
roofline.traffic):  python tools/make_traffic.py <rep> <key, e.g. c3/n1/tc> [note]"""
import csv, json, os, subprocess, sys

def main():
    rep, key = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, first = rows[0], rows[1], rows[2]
    def get(name):
        i = hdr.index(name)
        v = float(first[i].replace(",", ""))
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    dur = first[hdr.index("gpu__time_duration.sum")] + " " + units[hdr.index("gpu__time_duration.sum")]
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    d = json.load(open(path)) if os.path.exists(path) else {}
    d[key] = {"bytes": rd + wr, "dram_read": rd, "dram_write": wr, "kernel": first[hdr.index("Kernel Name")],
              "duration_under_ncu": dur, "source": os.path.basename(rep), "note": note}
    json.dump(d, open(path, "w"), indent=1, sort_keys=True)
    print(key, d[key])

if __name__ == "__main__":
    main()
