"""Summarise an .ncu-rep (full set, --import-source on) the way profiles/*_ncu_summary.md quote it:

    python tools/summarize_ncu.py gpurun_out/r02a_rhs_window.ncu-rep [--lines 25]

prints the headline metrics of every captured kernel, then stall samples and executed instructions per CUDA source line
(top N by samples) and per source file.  Runs here (no GPU): it only reads the report through `ncu -i`."""
import argparse
import collections
import csv
import io
import subprocess

METRICS = [
  'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
  'l1tex__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
  'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
  'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
  'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
  'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum',
  'l1tex__t_requests_pipe_lsu_mem_local_op_st.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
  'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
  'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
  'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
  'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
]


def ncu(report, *page):
  out = subprocess.run(['ncu', '-i', report, '--csv', *page], capture_output=True, text=True, check=True).stdout
  return list(csv.reader(io.StringIO(out)))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('report')
  ap.add_argument('--lines', type=int, default=25)
  args = ap.parse_args()
  rows = ncu(args.report, '--page', 'raw')
  hdr, units = rows[0], rows[1]
  for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')][:110])
    for m in METRICS:
      if m in hdr:
        print('   %-82s %s %s' % (m, r[hdr.index(m)], units[hdr.index(m)]))
  rows = ncu(args.report, '--page', 'source', '--print-source', 'cuda,sass')
  cur, hdr, per_line = None, None, []
  for r in rows:
    if len(r) == 2 and r[0] in ('File Path', 'File Name'):
      cur = r[1].split('/')[-1]
    elif r and r[0] == 'Line No':
      hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == '-':            # CUDA-source rows (SASS rows carry an address)
      try:
        per_line.append((cur, int(r[0]), r[1][:90], int(r[hdr.index('Instructions Executed')]),
                         int(r[hdr.index('# Samples')])))
      except ValueError:
        pass
  inst = sum(p[3] for p in per_line) or 1
  samp = sum(p[4] for p in per_line) or 1
  print('\n== top source lines by stall samples (%d samples, %d warp instructions)' % (samp, inst))
  for f, ln, src, i, s in sorted(per_line, key=lambda p: -p[4])[:args.lines]:
    print('   %-22s %4d  samples %5.1f%%  inst %5.1f%%  %s' % (f, ln, 100 * s / samp, 100 * i / inst, src))
  by_file = collections.Counter()
  for f, ln, src, i, s in per_line:
    by_file[f] += i
  print('\n== instructions by file')
  for f, i in by_file.most_common():
    print('   %-28s %5.1f%%' % (f, 100 * i / inst))


if __name__ == '__main__':
  main()
