# One-box A/B of the tensor-core step kernel: A = in-tree library (HEAD), B / C = variants under build_ab/ (see DESIGN 3.6 log).
set -x
mkdir -p gpurun_out
OUT=gpurun_out/r02_ab_tc.log
: > $OUT
( timeout 150 nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv -lms 500 > gpurun_out/r02_ab_clocks.csv 2>&1 ) &
A=tacotronv2_wavernn_chinese_b200/csrc/libb200tts.so
for pass in 1 2; do
  for v in A B C; do
    if [ $v = A ]; then L=$A; else L=build_ab/libb200tts_$v.so; fi
    echo "== variant $v pass $pass" >> $OUT
    ( B200TTS_LIB=$PWD/$L timeout 60 python tools/quick_time.py tc 256 3000 ) >> $OUT 2>&1
  done
done
cat $OUT
# winner: smallest best-of-two; its parity tests only if it is not A and at least 2 % faster
W=$(python - <<'PY'
import re
best = {}
v = None
for line in open('gpurun_out/r02_ab_tc.log'):
    m = re.match(r'== variant (\w) pass', line)
    if m: v = m.group(1)
    m = re.search(r'-> ([0-9.]+) us/step', line)
    if m and v: best[v] = min(best.get(v, 1e9), float(m.group(1)))
w = min(best, key=best.get) if best else 'A'
print(w if w != 'A' and best[w] < 0.98 * best.get('A', 0) else 'A')
PY
)
echo "winner $W" | tee -a $OUT
if [ "$W" != A ]; then
  ( B200TTS_LIB=$PWD/build_ab/libb200tts_$W.so timeout 150 python -m pytest tests/test_wavernn_gpu.py -q -x -k "test_tc_ or auto_dispatch" ) > gpurun_out/r02_ab_tests_$W.log 2>&1
  tail -3 gpurun_out/r02_ab_tests_$W.log
fi
