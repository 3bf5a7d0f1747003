#!/bin/bash
# One-GPU round check.  usage: bash tools/round_check.sh TAG [tests|bench|ncu]   (outputs under gpurun_out/, which may hold 64 MiB)
#   tests: the whole GPU test suite;  bench: the bench line (ours + reference arm) and the launch list of the bench;
#   ncu:   full captures of the kernels that carry the bench, condensed on the box (summary + hottest source lines); only the
#          headline report itself is kept
TAG=${1:-r2}; WHAT=${2:-all}
mkdir -p gpurun_out
if [ "$WHAT" = tests ] || [ "$WHAT" = all ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
fi
if [ "$WHAT" = bench ] || [ "$WHAT" = all ]; then
  timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 700 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
  timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference rc=$?"; tail -c 400 gpurun_out/${TAG}_bench_reference.json
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
  python tools/launch_summary.py gpurun_out/${TAG}_launches_bench.csv > gpurun_out/${TAG}_launches_summary.txt 2>&1
fi
if [ "$WHAT" = ncu ] || [ "$WHAT" = all ]; then
  cap() {  # name kernel-regex skip count command...
    local name=$1 kre=$2 skip=$3 cnt=$4; shift 4
    timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$kre" -s $skip -c $cnt -f -o gpurun_out/${TAG}_ncu_$name "$@" > gpurun_out/${TAG}_ncu_$name.log 2>&1
    echo "ncu $name rc=$?"
    python tools/ncu_summary.py gpurun_out/${TAG}_ncu_$name.ncu-rep gpurun_out/${TAG}_ncu_$name.txt > /dev/null 2>&1
    for k in $(echo $kre | tr '|' ' '); do
      echo "## hottest source lines of $k" >> gpurun_out/${TAG}_ncu_$name.txt
      python tools/ncu_source_lines.py gpurun_out/${TAG}_ncu_$name.ncu-rep $k 0 24 >> gpurun_out/${TAG}_ncu_$name.txt 2>&1
    done
  }
  cap headline k_fused2 3 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-adi
  cap row2 k_row2 2 1 python tools/row2_check.py
  rm -f gpurun_out/${TAG}_ncu_row2.ncu-rep
  cap fused2_mul k_fused2 2 1 python tools/prof_adi.py 8191
  rm -f gpurun_out/${TAG}_ncu_fused2_mul.ncu-rep
  POP_DD_FAKE_PEERS=1 cap dd8 "k_row2|k_fused2" 6 2 python tools/prof_dd.py 8 1 6
  rm -f gpurun_out/${TAG}_ncu_dd8.ncu-rep
fi
