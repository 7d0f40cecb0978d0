"""DRAM bytes per update of the captured launches of the two ncu reports (QLT, CAAS):
dram__bytes_read.sum + dram__bytes_write.sum per launch over the updates one launch
processes. Writes the JSON bench.py reads for roofline.traffic.
Usage: python tools/ncu_traffic.py qlt.ncu-rep caas.ncu-rep updates_per_launch out.json"""
import csv
import io
import json
import subprocess
import sys


def launches(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].split("<")[0].split("::")[-1].replace("void ", "")
        b = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[ix[m]].replace(",", ""))*scale[units[ix[m]]]
        out[name] = b
    return out


qrep, crep, upd, dst = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
res = {"source": "ncu --set full --clock-control none, tools/prof_run.py {qlt,caas} ne120x128x40 "
                 "1 640; dram__bytes_read.sum + dram__bytes_write.sum per launch, divided by the "
                 "%d updates of the launch" % int(upd),
       "updates_per_profiled_launch": int(upd), "bytes_per_update": {}}
for kind, rep in (("qlt", qrep), ("caas", crep)):
    d = {k: v/upd for k, v in launches(rep).items()}
    # run() only: the driver script's input generation and set_Qm are not part of it.
    d["run_total"] = sum(v for k, v in d.items()
                         if k not in ("fill_headline_kernel", "set_qm_bulk_kernel",
                                      "get_qm_bulk_kernel"))
    res["bytes_per_update"][kind] = d
json.dump(res, open(dst, "w"), indent=1)
print(json.dumps(res["bytes_per_update"], indent=1))
