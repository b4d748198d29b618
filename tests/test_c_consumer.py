"""The drop-in boundary from a compiled language: tests/c_consumer/consumer.c is a plain-C program (gcc, dlopen — what a Julia
`ccall` does) that describes a model as SoA iterator columns + postfix tapes through include/iexa.h, evaluates every
NLPModels callback with host pointers and checks the results against closed forms it computes itself.  No Python, torch or C++
on the consumer's side of the ABI.  Without a device the same program proves the evaluation entry points refuse to run."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_consumer", "consumer.c")


@pytest.fixture(scope="module")
def consumer(tmp_path_factory):
    import iexa_b200 as ex
    ex.lib.load()       # builds libiexa_b200.so if it is missing
    exe = str(tmp_path_factory.mktemp("c_consumer") / "consumer")
    subprocess.check_call(["gcc", "-std=gnu99", "-O1", "-Wall", "-Wextra", "-Werror", "-o", exe, SRC, "-ldl", "-lm"])
    so = os.path.join(ROOT, "infiniteexamodels.jl_b200", "libiexa_b200.so")
    assert os.path.exists(so)
    return exe, so


def test_plain_c_consumer_builds_a_plan_and_is_refused_evaluation_without_a_device(consumer):
    exe, so = consumer
    r = subprocess.run([exe, so, "nodevice"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("OK nodevice") and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_plain_c_consumer_evaluates_every_callback_on_the_gpu(consumer):
    exe, so = consumer
    r = subprocess.run([exe, so], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("OK:")
