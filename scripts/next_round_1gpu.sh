#!/bin/bash
# First 1-GPU call of the next round: full suite, then everything written after the round-1 GPU budget ran
# out (NOTES.md): DeviceTopK(pass1_fraction), the sketch backend="device" adapters, a bench sanity line.
#   gpurun --timeout 900 -- 'bash scripts/next_round_1gpu.sh'
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/ -x -q -m gpu > gpurun_out/nr_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 gpurun_out/nr_pytest_gpu.log
for f in 0.125 0.25; do
  timeout 120 python benchmarks/topk_bench.py --pass1-fraction $f > gpurun_out/nr_topk_f$f.json 2> gpurun_out/nr_topk_f$f.err
  python -c "import json;d=json.load(open('gpurun_out/nr_topk_f$f.json'));print('pass1_fraction',d['pass1_fraction'],'total',round(d['pass1_fraction_total_ms'],3),'ms vs plain',round(d['topk_total_ms'],3),'ms')" 2>/dev/null || tail -3 gpurun_out/nr_topk_f$f.err
done
timeout 120 python - <<'PY' 2>&1 | tail -5
import sys; sys.path.insert(0, ".")
from fps_b200.models.sketch.jobs import bloomFilter, tugOfWar, minhash
from fps_b200.models.sketch.device import bloom_positions64, tow_bits64, minhash_packed64
from fps_b200.models.sketch.hashing import java_string_hash
tweets = [(t, ["cat", "dog"] if t % 2 else ["cat", "car"]) for t in range(1, 40)]
b = dict(bloomFilter(tweets, 256, 3, 1, 1, backend="device"))
want = set()
for t, ws in tweets:
    if "cat" in ws: want |= set(bloom_positions64(t, 3, 256))
assert b[java_string_hash("cat")] == frozenset(want), "bloom adapter mismatch"
tw = dict(tugOfWar(tweets, 64, 1, 1, backend="device"))
assert tw[java_string_hash("cat")] == [sum(tow_bits64(t, 64)[j] for t, _ in tweets) for j in range(64)], "tow adapter mismatch"
mh = dict(minhash(tweets, 16, 1, 1, backend="device"))
exp = [min(minhash_packed64(t, 16)[j] for t, _ in tweets) & 0xFFFFFFFF for j in range(16)]
assert mh[java_string_hash("cat")] == exp, "minhash adapter mismatch"
print("SKETCH_DEVICE_ADAPTERS_OK")
PY
timeout 200 python bench.py --steps 100 --warmup 10 2> gpurun_out/nr_bench.err | tee gpurun_out/nr_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench N=1', round(d['value']/1e9,3), 'G/s e2e', round(d['e2e']['value']/1e9,3))"
