timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
DFB_GRU_NS=32 timeout 600 python -m pytest tests -m gpu -x -q -k "forward_random or golden or si_sdr" 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b.json 2>gpurun_out/b.err
python -c "import json; d=json.load(open('gpurun_out/b.json')); print('bench',d['value'], d['ms_per_step'], d['e2e']['value'], d['rtf_batch1']); print(d['roofline']['kernel_ms_per_step'])"
DFB_GRU_NS=16 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b16.json 2>gpurun_out/b.err
python -c "import json; d=json.load(open('gpurun_out/b16.json')); print('ns16',d['value'], d['ms_per_step']); print(d['roofline']['kernel_ms_per_step'])"
tail -3 gpurun_out/b.err
