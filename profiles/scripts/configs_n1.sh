#!/bin/bash
# BASELINE configs[2], [3], [4] through bench.py on one GPU
python bench.py --config 3 --steps 20 --warmup 3 > gpurun_out/r02_cfg3_n1.json 2> gpurun_out/r02_cfg3_n1.err
tail -c 400 gpurun_out/r02_cfg3_n1.err
for d in "1 1" "2 1" "2 2" "3 2" "3 3"; do
  t=$(echo $d | tr ' ' '_')
  python bench.py --config 5 --degree $d --steps 20 --warmup 3 --no-extras > gpurun_out/r02_cfg5_deg${t}_n1.json 2> gpurun_out/r02_cfg5.err
done
python bench.py --config 5 --degree 3 2 --steps 20 --warmup 3 > gpurun_out/r02_cfg5_deg3_2_n1_full.json 2>> gpurun_out/r02_cfg5.err
python bench.py --config 4 --steps 10 --warmup 3 > gpurun_out/r02_cfg4_n1.json 2> gpurun_out/r02_cfg4_n1.err
tail -c 600 gpurun_out/r02_cfg4_n1.err
python bench.py --config 4 --steps 10 --warmup 3 --backbone-amp --no-extras > gpurun_out/r02_cfg4_n1_amp.json 2>> gpurun_out/r02_cfg4_n1.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_cfg*_n1*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d.get("value", 0)), "e2e", round(d["e2e"]["value"]) if "e2e" in d else None, round(d.get("ms_per_step", 0), 3),
              "frac", d.get("roofline", {}).get("frac"), d.get("unavailable"))
    except Exception as e:
        print(f, "ERR", e)
PY
