"""Experiment: does running two half-batches on two streams hide the HBM-bound kernels behind the GEMMs?"""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import stlt_b200
from stlt_b200.synthetic import make_batch, random_state_dict

cfg = stlt_b200.StltModelConfig(num_classes=174, unique_categories=4)
torch.manual_seed(0)
sd = random_state_dict(stlt_b200.Stlt(cfg).state_dict(), seed=0)
models = []
for _ in range(3):
    m = stlt_b200.Stlt(cfg, precision=sys.argv[1] if len(sys.argv) > 1 else "bf16")
    m.load_state_dict(sd); m = m.to("cuda"); m.train(False); models.append(m)
full = {k: v.cuda() for k, v in make_batch(4096, "something", ragged=False, seed=1).items()}
halves = [{k: v[:2048].contiguous() for k, v in full.items()}, {k: v[2048:].contiguous() for k, v in full.items()}]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]

def run_single():
    models[2](full)

def run_dual():
    main = torch.cuda.current_stream()
    for i in range(2):
        streams[i].wait_stream(main)
        with torch.cuda.stream(streams[i]):
            models[i](halves[i])
    for i in range(2):
        main.wait_stream(streams[i])

def timeit(fn, n=10):
    with torch.no_grad():
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

for rep in range(2):
    print("single stream, B=4096: %.2f ms" % timeit(run_single))
    print("two streams, 2 x 2048: %.2f ms" % timeit(run_dual))
