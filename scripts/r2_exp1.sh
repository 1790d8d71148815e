set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -2; nproc
for cfg in "4096 4 64" "16384 2 32" "16384 3 48" "16384 4 64" "16384 8 64" "32768 3 24"; do
  set -- $cfg
  python bench.py --proofs $1 --lanes $2 --steps $3 --warmup 3 --extras 0 > gpurun_out/r2_exp1_p$1_l$2.json 2> gpurun_out/r2_exp1_p$1_l$2.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r2_exp1_p$1_l$2.json"))
    print("RESULT proofs=$1 lanes=$2 value=%.3e e2e=%.3e one=%.3e frac=%.3f" % (d["value"], d["e2e"]["value"], d["one_batch_at_a_time"]["value"], d["roofline"]["whole_step"]["frac"]))
    print({k:round(v["ms"],4) for k,v in d["roofline"]["per_kernel"].items()})
    print(d["e2e"]["one_call_at_a_time"])
except Exception as e: print("ERR", e)
P
done
