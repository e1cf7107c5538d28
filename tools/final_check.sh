#!/bin/bash
# one-call end-of-session check: GPU test suite, smoke(), the default bench line, A/B of the channel splits
tag=${1:-final}
o=gpurun_out/${tag}
timeout 150 python -m pytest tests -m gpu -x -q > ${o}_pytest.log 2>&1; rc=$?; echo "rc=$rc" >> ${o}_pytest.log
tail -3 ${o}_pytest.log
extra=""
if [ $rc -ne 0 ]; then
    DAVI_PLAIN_SLICES=1 timeout 150 python -m pytest tests -m gpu -x -q > ${o}_pytest_plain.log 2>&1; echo "rc=$?" >> ${o}_pytest_plain.log
    tail -3 ${o}_pytest_plain.log
    extra="--plain-slices"
fi
timeout 90 python __graft_entry__.py smoke > ${o}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 ${o}_smoke.log
timeout 60 python bench.py --steps 8 --warmup 3 --no-roofline --no-cpu-baseline --plain-slices > ${o}_bench_plain.json 2> ${o}_bench_plain.err
timeout 400 python bench.py $extra > ${o}_bench.json 2> ${o}_bench.err; echo "bench rc=$?"
python - <<PY
import json
for f in ("bench_plain", "bench"):
    try:
        d = json.loads(open("${o}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), round(d["ms_per_step"], 2), d["e2e"]["value"], d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 300 ${o}_bench.err
