mkdir -p gpurun_out/scale
python bench.py --gpus 1 --steps 5 --warmup 3 --skip-cpu --skip-extras > gpurun_out/scale/n1.json 2> gpurun_out/scale/n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 5 --warmup 3 --skip-cpu --skip-extras > gpurun_out/scale/n$n.json 2> gpurun_out/scale/n$n.err
done
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.load(open(f'gpurun_out/scale/n{n}.json'))
        print(n, round(d['value'],3), round(d['ms_per_step'],3), round(d['e2e']['value'],3), round(d['sweep_1024x2048']['ms'],3), d['sweep_1024x2048']['nlml_checksum'])
    except Exception as e:
        print(n, 'failed', e)
PY
