#!/usr/bin/env python
"""Turns an `ncu --page raw --csv` dump of one captured kernel into profiles/r02_ncu_traffic.json (what bench.py reports as
roofline.traffic).  Usage: ncu -i cap.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_traffic.py raw.csv agg_update_kernel <rows> <source>"""
import csv
import json
import os
import sys


def main():
    path, kernel, rows, source = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    rd = list(csv.reader(open(path)))
    hdr = rd[0]
    units = rd[1]
    row = rd[2]
    def metric(name):
        i = hdr.index(name)
        v = float(row[i].replace(",", ""))
        u = units[i].lower()
        mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[u]
        return v * mult
    total = metric("dram__bytes_read.sum") + metric("dram__bytes_write.sum")
    dur_i = hdr.index("gpu__time_duration.sum")
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_ncu_traffic.json")
    data = json.load(open(out)) if os.path.exists(out) else {}
    data[kernel] = {"rows": rows, "dram_bytes": int(total), "duration": "%s %s" % (row[dur_i], units[dur_i]), "source": source}
    json.dump(data, open(out, "w"), indent=1)
    print(data[kernel])


if __name__ == "__main__":
    main()
