python - <<'PY'
import torch, time
x = torch.empty(502_000_000, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); d.copy_(x, non_blocking=True); torch.cuda.synchronize(); t = time.perf_counter() - t0
print("H2D 502 MB pinned: %.2f ms (%.1f GB/s)" % (t * 1e3, 0.502 / t))
PY
nproc; lscpu | grep "Model name"
for cfg in "4 4" "4 5" "5 5" "6 6" "8 6" "8 8" "12 8"; do set -- $cfg; python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-stage5 --chunks $1 --streams $2 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('chunks $1 streams $2', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2))"; done
