"""The reference's own benchmark driver, UNMODIFIED (/root/reference/benchmark.py, or its verbatim git-ignored copy under
baseline/_ref on the GPU box), as the external clock (BASELINE.json configs[0]):

  * CPU: `benchmark.py --model lemevit_tiny --bench inference -b 1 --img-size 224 --device cpu` with the reference's own `models`
    package — the plumbing run of BASELINE config 1 (no GPU);
  * GPU: the same file with `tests/drop_in/models` (= `from lemevit_b200 import *`) in place of the reference's `models` package:
    `--model lemevit_base --precision bfloat16 -b 256` on a B200 — the drop-in claim, measured by the reference's clock.

timm is not installed; tests/timm_shim provides the symbols the driver imports (tests/timm_shim/README.md)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from oracle import ref_copy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "timm_shim")
DROP_IN = os.path.join(ROOT, "tests", "drop_in")
REF = ref_copy.root()
needs_ref = pytest.mark.skipif(REF is None, reason="reference sources not available (neither /root/reference nor baseline/_ref)")


def run_driver(pythonpath, args, cwd, timeout=900):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(pythonpath)
    env.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
    # run the file through runpy so that its own directory (which holds the reference's `models` package) is NOT put in front of
    # PYTHONPATH: which `models` package `from models import *` (benchmark.py:70) finds is decided by `pythonpath` alone
    script = os.path.join(REF, "benchmark.py")
    boot = "import runpy, sys; sys.argv = [sys.argv[1]] + sys.argv[2:]; runpy.run_path(sys.argv[0], run_name='__main__')"
    proc = subprocess.run([sys.executable, "-c", boot, script] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=timeout)
    out = proc.stdout + proc.stderr
    assert proc.returncode == 0, out[-4000:]
    assert "--result" in proc.stdout, out[-4000:]
    return json.loads(proc.stdout.split("--result", 1)[1]), out


@needs_ref
def test_unmodified_benchmark_py_cpu_plumbing(tmp_path):
    """BASELINE.json configs[0]: LeMeViT-Tiny 224x224 batch 1 forward on CPU via benchmark.py, reference model, no GPU."""
    res, out = run_driver([SHIM, REF], ["--model", "lemevit_tiny", "--bench", "inference", "-b", "1", "--img-size", "224", "--device", "cpu",
                                        "--num-warm-iter", "2", "--num-bench-iter", "5"], cwd=str(tmp_path))
    assert res["model"] == "lemevit_tiny" and "error" not in res, res
    assert res["param_count"] == 8.64                     # README.md:85
    assert res["infer_batch_size"] == 1 and res["infer_img_size"] == 224 and res["infer_samples_per_sec"] > 0


@needs_ref
def test_drop_in_models_package_registers_the_native_entrypoints(tmp_path):
    """`from models import *` (benchmark.py:70) with tests/drop_in first on the path resolves to lemevit_b200 and registers with timm;
    on a CPU-only machine the driver then fails LOUDLY at the first forward (no CPU fallback) — checked through the driver's own
    error channel."""
    env_path = [SHIM, DROP_IN, ROOT]
    code = ("import timm.models as T, models, lemevit_b200;"
            "assert T.is_model('lemevit_base') and T.model_entrypoint('lemevit_base') is lemevit_b200.lemevit_base;"
            "m = T.create_model('lemevit_tiny', num_classes=None, in_chans=3, global_pool=None, scriptable=False, drop_rate=0.0, drop_path_rate=None);"
            "assert type(m).__module__ == 'lemevit_b200.model' and m.default_cfg['input_size'] == (3, 224, 224); print('ok')")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(env_path))
    proc = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0 and "ok" in proc.stdout, proc.stdout + proc.stderr
    if not torch.cuda.is_available():
        res, out = run_driver(env_path, ["--model", "lemevit_tiny", "--bench", "inference", "-b", "1", "--device", "cpu", "--num-warm-iter", "1",
                                         "--num-bench-iter", "5", "--no-retry"], cwd=str(tmp_path))
        assert "error" in res and "no CPU fallback" in res["error"], res


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("model,batch,floor", [("lemevit_base", 256, 6000.0), ("lemevit_tiny", 256, 15000.0)])
def test_unmodified_benchmark_py_drives_the_native_model_on_gpu(tmp_path, model, batch, floor):
    """benchmark.py:409-431,481-519 unchanged, bf16, batch 256: model built by timm's create_model from the native entrypoints,
    `.to(device, dtype)`, `.eval()`, 10 warm-up + 40 timed steps with a synchronize per step.  The img/s of the reference's own
    clock is logged to gpurun_out/ (copied to profiles/ by the round script)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    res, out = run_driver([SHIM, DROP_IN, ROOT], ["--model", model, "--bench", "inference", "-b", str(batch), "--img-size", "224",
                                                  "--precision", "bfloat16", "--num-bench-iter", "40", "--no-retry"], cwd=str(tmp_path))
    assert "error" not in res, res
    assert res["infer_batch_size"] == batch and res["infer_samples_per_sec"] > floor, res
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"reference_driver_{model}_b{batch}_native.json"), "w") as f:
        json.dump({"driver": "unmodified reference benchmark.py", "models_package": "tests/drop_in/models (from lemevit_b200 import *)",
                   "args": f"--model {model} --bench inference -b {batch} --img-size 224 --precision bfloat16 --num-bench-iter 40", "result": res}, f, indent=1)
