"""tests/golden/lr_schedule.json from the UNMODIFIED reference scheduler (authoring container; TEST INFRASTRUCTURE).
Steps utils/schedulers.py:LinearWarmupCosineAnnealingLR once per epoch exactly like Lightning does for train.py:71-84."""
import importlib.util, json, os, sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("ref_schedulers", "/root/reference/utils/schedulers.py")
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)
out = {}
for name, (lr, epochs) in {"train_default": (2e-4, 500), "short": (1e-3, 30), "no_warmup_5": (2e-4, 5), "no_warmup_9": (2e-4, 9)}.items():
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=lr)
    sch = mod.LinearWarmupCosineAnnealingLR(optimizer=opt, warmup_epochs=int(0.1 * epochs), max_epochs=epochs, eta_min=1e-6)
    lrs = []
    for _ in range(epochs):
        lrs.append(opt.param_groups[0]["lr"])   # lr in effect during this epoch
        opt.step()
        sch.step()
    out[name] = {"base_lr": lr, "max_epochs": epochs, "warmup_epochs": int(0.1 * epochs), "eta_min": 1e-6, "lrs": lrs}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "lr_schedule.json"), "w"))
print({k: (v["lrs"][:3], v["lrs"][-1]) for k, v in out.items()})
