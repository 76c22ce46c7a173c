#!/bin/bash
# BASELINE configs 3 (single-level LS), 4 (VGG) and 5 (eval) on N GPUs of one box; $1 = N, $2 = tag
N=${1:-1}; TAG=${2:-r02}
mkdir -p gpurun_out
run() {  # name, bench args...
  name=$1; shift
  if [ "$N" -gt 1 ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N "$@" > gpurun_out/bench_${name}_${N}gpu_$TAG.json 2> gpurun_out/bench_${name}_${N}gpu_$TAG.err
  else
    timeout 900 python bench.py "$@" > gpurun_out/bench_${name}_${N}gpu_$TAG.json 2> gpurun_out/bench_${name}_${N}gpu_$TAG.err
  fi
  echo "== $name N=$N exit $?"; tail -n 2 gpurun_out/bench_${name}_${N}gpu_$TAG.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${name}_${N}gpu_$TAG.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("metric","value","n_gpus","ms_per_step")}, "e2e", d["e2e"]["value"], "cpu", (d.get("cpu_baseline") or {}).get("value"), "gpu_ref", {k:v for k,v in (d.get("reference_gpu_eager") or {}).items() if k.startswith("step_ms")})
except Exception as e: print("no line", e)
PY
}
EXTRA=""
if [ "$N" -gt 1 ]; then EXTRA="--no-cpu-baseline"; fi
WHICH=${3:-multi,singleLS,vgg,eval}
case ",$WHICH," in *,multi,*) run multi --steps 20 --warmup 5 --also-trunk-bf16 0 $EXTRA;; esac
case ",$WHICH," in *,singleLS,*) run singleLS --level single-level --gan LS --steps 20 --warmup 5 --also-trunk-bf16 0 $EXTRA;; esac
case ",$WHICH," in *,vgg,*) run vgg --model VGG --gan LS --steps 20 --warmup 5 --also-trunk-bf16 0 $EXTRA;; esac
case ",$WHICH," in *,eval,*) run eval --mode eval --frames 500 $EXTRA;; esac
