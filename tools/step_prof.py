"""One profiled Bloom-560M training step for ncu (--profile-from-start off): 2 warm-up steps, then
cudaProfilerStart / one step / cudaProfilerStop. Usage under gpurun:
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python tools/step_prof.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cleantransformer_b200.models import modeling_bloom as mb
from cleantransformer_b200.optimizer import TorchAdamW

layers = int(os.environ.get("LAYERS", "24"))
cfg = dict(vocab_size=250880, hidden_size=1024, n_layer=layers, num_attention_heads=16)
dev = torch.device("cuda", 0)
torch.manual_seed(999)
with torch.device(dev):
    model = mb.BloomForCausalLM(mb.BloomConfig(**cfg))
with torch.no_grad():
    for _, p in model.named_parameters():
        if p.dim() >= 2:
            p.normal_(0.0, 0.02)
model._tie_weight(); model.train()
opt = TorchAdamW(model.parameters(), lr=1e-5)
ids = torch.randint(3, 250880, (8, 1024), device=dev); mask = torch.ones(8, 1024, dtype=torch.long, device=dev)

def step():
    opt.zero_grad()
    out, _ = model(input_ids=ids, attention_mask=mask, labels=ids)
    out[0].backward()
    opt.step()

for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step")
