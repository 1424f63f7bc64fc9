"""Print per-parameter gradient errors of the CUDA backward vs the oracle's autograd (debug aid, GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_train_model_gpu import run_backward, grad_errors
from tests.train_helpers import oracle_grads
from tests.conftest import rel_err

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
ref_out, ref_grads = oracle_grads()
net, tr, out = run_backward(prec)
print("forward rel err", rel_err(out.cpu(), ref_out))
errs = grad_errors(net, ref_grads)
for n, e in sorted(errs.items(), key=lambda t: -t[1] if t[1] == t[1] else -1e9)[:60]:
    print(f"{e:10.3e}  {n}")
import statistics
vals = [e for e in errs.values() if e == e]
print("n", len(errs), "nan", len(errs) - len(vals), "median", statistics.median(vals), "max", max(vals))
