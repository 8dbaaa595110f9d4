"""bench.py without a GPU: every global name its functions load is defined somewhere in the module (the multi-GPU branch only
runs under torchrun on a GPU box, so a lost module-level constant would otherwise surface there), and the lane policy
(lane_ctas: how the batches in flight share the SMs through gomp_batch_set_max_ctas)."""
import ast
import builtins
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_has_no_undefined_names():
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    defined = set(dir(builtins)) | {"__file__", "__name__"}
    for n in ast.walk(tree):
        if isinstance(n, (ast.FunctionDef, ast.ClassDef)):
            defined.add(n.name)
        elif isinstance(n, ast.Import):
            defined.update((a.asname or a.name).split(".")[0] for a in n.names)
        elif isinstance(n, ast.ImportFrom):
            defined.update(a.asname or a.name for a in n.names)
        elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            defined.add(n.id)
        elif isinstance(n, ast.arg):
            defined.add(n.arg)
        elif isinstance(n, ast.ExceptHandler) and n.name:
            defined.add(n.name)
    missing = sorted({n.id for n in ast.walk(tree) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in defined})
    assert not missing, missing


def test_lane_policy():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    L, sms = bench.LANES, 148
    assert L >= 1
    # constant share: every step gets num_sms / LANES CTAs (rounded up), one lane = no cap
    caps = [bench.lane_ctas(k, 20, sms, 0) for k in range(20)]
    assert len(set(caps)) == 1 and (caps[0] == 0 if L == 1 else caps[0] * L >= sms > (caps[0] - 1) * L)
    # ramp at both ends: the first and the last batch of a run are uncapped, the middle ones get the constant share, symmetric
    caps = [bench.lane_ctas(k, 20, sms, 3) for k in range(20)]
    assert caps[0] == 0 and caps[-1] == 0 and caps == caps[::-1]
    assert all(c == bench.lane_ctas(5, 20, sms, 0) for c in caps[L - 1:20 - L + 1])
    if L > 2:
        assert caps[1] == -(-sms // 2)
    # a single step is never capped
    assert bench.lane_ctas(0, 1, sms, 0) == (0 if L == 1 else -(-sms // L)) and bench.lane_ctas(0, 1, sms, 3) == 0
