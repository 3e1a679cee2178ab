"""Summarise an .ncu-rep (raw page + stall breakdown from the source page) into a text file."""
import csv, io, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.avg.per_cycle_active']


def main(rep, out=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        lines.append("kernel: %s  grid %s block %s" % (d.get('Kernel Name'), d.get('Grid Size'), d.get('Block Size')))
        for k in KEYS:
            if k in d:
                lines.append("  %-78s %s %s" % (k, d[k], units[hdr.index(k)]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = None
    for i, r in enumerate(rows):
        if r and r[0] == "Address":
            h = r; body = rows[i + 1:]; break
    if h:
        idx = {n: i for i, n in enumerate(h)}
        stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        tot = {s: 0 for s in stalls}; total = 0; byop = {}
        for r in body:
            if len(r) < len(h):
                continue
            n = int(r[idx['# Samples']] or 0); total += n
            for s in stalls:
                tot[s] += int(r[idx[s]] or 0)
            toks = r[idx['Source']].split()
            op = (toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?'))
            byop[op] = byop.get(op, 0) + n
        lines.append("  warp-stall samples (all): total %d" % total)
        for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
            lines.append("    %-26s %6.3f" % (s, v / max(total, 1)))
        lines.append("  samples by opcode:")
        for o, v in sorted(byop.items(), key=lambda kv: -kv[1])[:12]:
            lines.append("    %-26s %6.3f" % (o, v / max(total, 1)))
    txt = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
