"""Summarise ncu outputs (read here, on the CPU box): launch-list CSV -> per-kernel shares; .ncu-rep -> key raw metrics."""
import collections, csv, subprocess, sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'smsp__inst_executed.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_bytes.sum',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum']


def launches(path, header):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]
    ci = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            t = float(r[ci['Metric Value']])
        except ValueError:
            continue
        u = r[ci['Metric Unit']]
        t = t / 1000.0 if u in ('ns', 'nsecond') else (t * 1000.0 if u in ('ms', 'msecond') else t)
        a = agg.setdefault(r[ci['Kernel Name']], [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    out = list(header) + ["# kernel | launches | total us | share | us per launch"]
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append("%s | %d | %.1f | %.3f | %.1f" % (k[:150], v[0], v[1], v[1] / tot, v[1] / v[0]))
    return "\n".join(out) + "\n"


def raw(rep, header):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    out = list(header)
    for i, h in enumerate(hdr):
        if h == 'Kernel Name' or h in KEEP:
            out.append("%s [%s]: %s" % (h, rows[1][i], " | ".join(r[i] for r in rows[2:])))
    return "\n".join(out) + "\n"


def stalls(rep, kernel_index=0, top=14):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    blocks, cur = [], []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            blocks.append(cur); cur = []; continue
        if r and r[0] == "Address":
            continue
        cur.append(r)
    blocks.append(cur)
    data = blocks[kernel_index]
    si = ci['# Samples']
    tot = sum(float(r[si]) for r in data) or 1.0
    st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = sorted(((sum(float(r[ci[h]]) for r in data), h) for h in st), reverse=True)[:6]
    out = ["# warp-state samples: total %d; by reason: %s" % (tot, ", ".join("%s %.0f%%" % (h, 100 * v / tot) for v, h in agg)),
           "# top sampled SASS lines: samples | share | instruction | top stall"]
    for r in sorted(data, key=lambda r: -float(r[si]))[:top]:
        s2 = sorted(((float(r[ci[h]]), h) for h in st), reverse=True)[0]
        out.append("%d | %.1f%% | %s | %s" % (float(r[si]), 100 * float(r[si]) / tot, r[1].strip()[:80], s2[1]))
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    kind = sys.argv[1]
    if kind == "launches":
        sys.stdout.write(launches(sys.argv[2], sys.argv[3:]))
    elif kind == "raw":
        sys.stdout.write(raw(sys.argv[2], sys.argv[3:]))
        sys.stdout.write(stalls(sys.argv[2]))
