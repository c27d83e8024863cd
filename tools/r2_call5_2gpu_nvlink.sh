#!/bin/bash
# Round 2, GPU call 5 (2 GPUs, short): ncu NVLink byte counters of the peer-read and peer-store kernels (one process drives
# both GPUs, tools/rowmove_lab.cu).  The profiler serialises kernels, so the one-directional numbers are also true rates.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=wholegraph_b200/lib/rowmove_lab
M=gpu__time_duration.sum
for d in rx tx; do for m in bytes bytes_data_user bytes_data_protocol bytes_packet_request bytes_packet_response bytes_packet_request_data_protocol bytes_packet_request_data_user bytes_packet_response_data_protocol bytes_packet_response_data_user; do M=$M,nvl${d}__$m.sum; done; done
M=$M,nvlrx__cycles_active.avg,nvlrx__cycles_elapsed.avg,nvltx__cycles_active.avg,nvltx__cycles_elapsed.avg
for what in "uni gather" "uni scatter" "bidir gather" "bidir scatter"; do
  set -- $what
  out=gpurun_out/r2_nvlink_$1_$2_rb1024
  timeout 300 ncu --metrics $M --clock-control none -k regex:row_move_vec -s 8 -c 2 --csv --log-file $out.csv $L --row-bytes 1024 --rows 20000000 --mode $1 --op $2 --set default --iters 4 --warmup 4 > $out.log 2>&1
  echo "== $what"; python - $out.csv <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
i_id, i_name, i_metric, i_unit, i_val = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
i_dev = hdr.index("Device") if "Device" in hdr else None
seen = {}
for r in rows[1:]:
    seen.setdefault((r[i_id], r[i_dev] if i_dev is not None else ""), {})[r[i_metric]] = (r[i_val], r[i_unit])
for (kid, dev), m in seen.items():
    print("launch", kid, "device", dev)
    for k in sorted(m):
        print("   %-52s %18s %s" % (k, m[k][0], m[k][1]))
PY
done
