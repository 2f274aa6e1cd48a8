"""Developer tool: turn ncu output into the text summaries kept under profiles/.

  python tools/ncu_summary.py launches <ncu --csv log> "<header line>"      per-kernel totals / shares of a launch list
  python tools/ncu_summary.py traffic <file.ncu-rep> "<header line>"        DRAM traffic + tensor-pipe activity of a --set full capture
"""
import csv, subprocess, sys, io, collections


def launches(path, header):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    h = rows[0]
    ki, mi, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    ui = h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        name = r[ki].split("(")[0].replace("void ", "").replace("fdbm::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(header)
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:52]:52s} n={n:5d} total={ms:9.3f} ms share={ms / tot:6.3f} avg={ms / n * 1e3:8.1f} us")
    print(f"{'TOTAL':52s} n={sum(a[0] for a in agg.values()):5d} total={tot:9.3f} ms")


def traffic(rep, header):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units, data = rows[0], rows[1], rows[2:]
    col = {c: i for i, c in enumerate(h)}

    def get(r, name, scale_unit=None):
        v = float(r[col[name]].replace(",", ""))
        u = units[col[name]]
        if scale_unit == "bytes":
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        if scale_unit == "ms":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1)
        return v
    rd = [get(r, "dram__bytes_read.sum", "bytes") for r in data]
    wr = [get(r, "dram__bytes_write.sum", "bytes") for r in data]
    ms = [get(r, "gpu__time_duration.sum", "ms") for r in data]
    tname = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    tn = [get(r, tname) for r in data] if tname in col else [0.0] * len(data)
    print(header)
    print(f"launches                      {len(data)}")
    print(f"dram__bytes_read.sum  total   {sum(rd) / 1e9:.3f} GB")
    print(f"dram__bytes_write.sum total   {sum(wr) / 1e9:.3f} GB")
    print(f"dram bytes per launch (avg)   {(sum(rd) + sum(wr)) / len(data) / 1e6:.1f} MB")
    print(f"gpu__time_duration.sum total  {sum(ms):.3f} ms (under ncu: cold caches, one launch at a time)")
    print(f"{tname}, time-weighted over the launches: {sum(a * b for a, b in zip(tn, ms)) / sum(ms):.1f} %")
    print("top 8 launches by time: (ms, tensor-pipe active %, dram read MB, dram write MB)")
    for i in sorted(range(len(data)), key=lambda i: -ms[i])[:8]:
        print(f"   {ms[i]:6.3f} ms  {tn[i]:5.1f} %  {rd[i] / 1e6:8.1f}  {wr[i] / 1e6:8.1f}")


def traffic_csv(path, micro_batch, n_frames, header=""):
    """`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active... --csv
    --log-file <path>` over the conv_igemm launches of ONE forward -> text summary on stdout + profiles/conv_traffic.json, the file
    bench.py reads `roofline.traffic` from (it carries the sha256 of the conv_igemm.cu it was captured on)."""
    import hashlib, json, os
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    h = rows[0]
    ii, mi, vi, ui = h.index("ID"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    per = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        if "byte" in u:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        elif r[mi].startswith("gpu__time"):
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        per.setdefault(r[ii], {})[r[mi]] = v
    L = list(per.values())
    rd = sum(d.get("dram__bytes_read.sum", 0) for d in L); wr = sum(d.get("dram__bytes_write.sum", 0) for d in L)
    ms = sum(d.get("gpu__time_duration.sum", 0) for d in L)
    tname = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    tw = sum(d.get(tname, 0) * d.get("gpu__time_duration.sum", 0) for d in L) / max(ms, 1e-9)
    print(header)
    print(f"launches {len(L)}; dram read {rd / 1e9:.3f} GB, written {wr / 1e9:.3f} GB, per launch {(rd + wr) / len(L) / 1e6:.1f} MB; "
          f"time under ncu {ms:.3f} ms; tensor pipe active, time-weighted {tw:.1f} %")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "rethinking-flow-and-diffusion-bridge-models-for-speech-enhancement_b200", "csrc", "conv_igemm.cu"), "rb").read()
    out = os.path.join(root, "profiles", "conv_traffic.json")
    rec = {"conv_igemm_sha256": hashlib.sha256(src).hexdigest(), "source": "ncu metrics pass over the conv_igemm launches of one forward (" + os.path.basename(path) + ")",
           "captures": [{"micro_batch": int(micro_batch), "n_frames": int(n_frames), "launches": len(L), "dram_bytes": rd + wr,
                         "dram_read_bytes": rd, "dram_write_bytes": wr, "tensor_pipe_active_pct_time_weighted": tw}]}
    json.dump(rec, open(out, "w"), indent=1)
    print("wrote", out)


if __name__ == "__main__":
    if sys.argv[1] == "traffic_csv":
        traffic_csv(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else "")
    else:
        {"launches": launches, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
