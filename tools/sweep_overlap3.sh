B="python bench.py --steps 6 --warmup 3 --no-extra --no-e2e --no-cpu --no-parity"
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
print("RESULT", sys.argv[2], "|", round(d["value"]), round(r["whole_step"]["ms_per_clip"],4), "serial", round(r["serial_step"]["ms"],4), {k:round(v["ms"],3) for k,v in r["stages"].items()}, d["clocks"]["sm_mhz"])
PY
}
i=0
while read -r knobs; do
  i=$((i+1))
  env $(echo "$knobs" | tr ' ' '\n' | grep = | tr '\n' ' ') $B $(echo "$knobs" | tr ' ' '\n' | grep -v = | tr '\n' ' ') > /tmp/s$i.json 2>/dev/null
  show /tmp/s$i.json "$knobs"
done
