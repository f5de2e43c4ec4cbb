"""The drop-in packages overlay a checkout of the reference: sub-modules we do not provide resolve to the
reference's files, and the reference's own modules bind OUR hot-path classes.  Needs /root/reference (build
container only) -> skipped elsewhere.  Runs in a subprocess so the package caches of this process stay clean."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PDM_REFERENCE_ROOT", "/root/reference")

SCRIPT = r'''
import os, sys
sys.path.insert(0, os.path.join(ROOT, "physics-of-diffusion-models_b200"))
sys.path.insert(0, ROOT)
from oracle.ref_loader import _install_stubs
_install_stubs()                       # diffusers / torchmetrics / ... are not installed in this image
import utils, diffusion
from diffusion.scheduler import Scheduler, CosineScheduler, LinearBetaScheduler
assert utils.__file__.startswith(os.path.join(ROOT, "physics-of-diffusion-models_b200"))
assert CosineScheduler.__module__ == "diffusion.scheduler.cosine"
import diffusion.scheduler.cosine as cos
assert cos.__file__.startswith(REF), cos.__file__
assert issubclass(CosineScheduler, Scheduler)
assert Scheduler.true_posterior_mean_x0.__module__ == "diffusion.scheduler.scheduler"
import diffusion.scheduler.scheduler as ours
assert ours.__file__.startswith(os.path.join(ROOT, "physics-of-diffusion-models_b200"))
# a helper that only the reference provides, reached through our package
assert utils.get_default_device() in ("cuda", "cpu", "mps")
assert utils.dict_map(lambda v: v + 1, {"a": 1}) == {"a": 2}
import utils.stats
assert utils.stats.__file__.startswith(os.path.join(ROOT, "physics-of-diffusion-models_b200"))
from diffusion import DDPMSampler      # the reference's sampler, importing OUR DDPM / Scheduler
import diffusion.ddpm_sampling as samp
assert samp.__file__.startswith(REF) and samp.DDPM is diffusion.DDPM
print("overlay ok")
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "utils")), reason="reference checkout not present")
def test_overlay_on_reference_checkout():
    env = dict(os.environ, PDM_REFERENCE_ROOT=REF)
    code = f"ROOT = {ROOT!r}\nREF = {REF!r}\n" + SCRIPT
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300, cwd=REF)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "overlay ok" in r.stdout
