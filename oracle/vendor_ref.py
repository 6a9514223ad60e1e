"""TEST / BASELINE INFRASTRUCTURE ONLY -- "install" the UNMODIFIED reference into the git-ignored `baseline/_ref/`.

The reference (zcaicaros/L2S) is a script collection without setup.py / pyproject.toml, so the offline
`pip install --target baseline/_ref /root/reference` of the bench contract does not apply; its install is a
verbatim file copy of the modules on (or next to) the hot path plus the two small data sets and weights the
BASELINE configs name.  `baseline/_ref/` is listed in .gitignore (no reference source enters the history) but
NOT in .gpurunignore, so it travels to the GPU box, where `/root/reference` does not exist:

  * `bench.py --impl reference` times the real `JsspN5.step` from there on the box's host cores;
  * `tests/test_gpu_live_reference.py` compares the CUDA path with the live reference on the box;
  * `tests/test_gpu_policy.py` runs the reference's own `model/actor.py` against `l2s_b200.policy.Actor`.

    python -m oracle.vendor_ref            # copy if /root/reference is present (idempotent)
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("L2S_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")

MODEL = "incumbent_{s}[1,99]_fdd-divide-mwkr_yaoxin_1_128_4_4_gin+dghan_1_0.0_5e-05_10_500_64_128000_10.pth"
FILES = [
    "__init__.py", "parameters.py", "conventional_baselines_with_long-term-mem.py", "test_learned.py",
    "n-step_reinforce.py", "ortools_solver.py",
    "env/__init__.py", "env/environment.py", "env/message_passing_evl.py", "env/permissible_LS.py",
    "env/generateJSP.py",
    "model/__init__.py", "model/actor.py",
    "test_data/syn10x10.npy", "test_data/syn10x10_result.npy", "test_data/tai20x15.npy", "test_data/tai20x15_result.npy",
    "validation_data/validation_instance_10x10[1,99].npy", "validation_data/validation10x10_ortools_result.npy",
    "saved_model/" + MODEL.format(s="10x10"), "saved_model/" + MODEL.format(s="20x15"),
]


def available():
    return os.path.isfile(os.path.join(DST, "env", "environment.py"))


def vendor(verbose=False):
    """Copy FILES from the reference tree (if present).  Returns the destination, or None if there is no source."""
    if not os.path.isfile(os.path.join(SRC, "env", "environment.py")):
        return DST if available() else None
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.isfile(s):
            if verbose:
                print("vendor_ref: missing in the reference:", rel)
            continue
        if os.path.isfile(d) and os.path.getsize(d) == os.path.getsize(s) and os.path.getmtime(d) >= os.path.getmtime(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        if verbose:
            print("vendor_ref:", rel)
    return DST


if __name__ == "__main__":
    print(vendor(verbose=True))
    sys.exit(0)
