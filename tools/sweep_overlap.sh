set -x
B="python bench.py --steps 6 --warmup 3 --no-extra --no-e2e --no-cpu --no-parity"
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
print("RESULT", round(d["value"]), round(r["whole_step"]["ms_per_clip"],4), "serial", round(r["serial_step"]["ms"],4), {k:round(v["ms"],3) for k,v in r["stages"].items()}, d["clocks"]["sm_mhz"])
PY
}
$B > /tmp/a.json 2>/dev/null; show /tmp/a.json
ELVIS_UMMA_SMEM_PAD=50000 $B > /tmp/b.json 2>/dev/null; show /tmp/b.json
ELVIS_UMMA_SMEM_PAD=50000 $B --move-ctas 4 > /tmp/c.json 2>/dev/null; show /tmp/c.json
ELVIS_UMMA_SMEM_PAD=50000 $B --move-ctas 2 > /tmp/d.json 2>/dev/null; show /tmp/d.json
ELVIS_UMMA_SMEM_PAD=50000 $B --depth 4 > /tmp/e.json 2>/dev/null; show /tmp/e.json
ELVIS_UMMA_SMEM_PAD=50000 ELVIS_PIPE_PRIO=0,0,0 $B > /tmp/f.json 2>/dev/null; show /tmp/f.json
ELVIS_UMMA_SMEM_PAD=50000 ELVIS_PIPE_PRIO=-1,0,0 $B > /tmp/g.json 2>/dev/null; show /tmp/g.json
