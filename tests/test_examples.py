"""The shipped example task files validate, and the CPU one runs to completion through the CLI."""
import json
import os
import subprocess
import sys

from olearning_sim_b200.api.schema import json2taskconfig
from olearning_sim_b200.api.validate import ValidateParameters
from olearning_sim_b200.deviceflow.strategy import Strategy
from olearning_sim_b200.deviceflow.validate import ValidateParameters as FlowValidate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    with open(os.path.join(ROOT, "examples", name)) as f:
        return json.load(f)


def test_examples_pass_the_three_validation_passes():
    for name in ("task_mlp_cpu.json", "task_cnn2_gpu.json"):
        ok, msg = ValidateParameters().validate_task_parameters(json2taskconfig(_load(name)))
        assert ok, (name, msg)


def test_gpu_example_strategy_is_a_valid_rate_curve():
    op = _load("task_cnn2_gpu.json")["operatorflow"]["operators"][0]
    st = json.loads(op["operation_behavior_controller"]["strategy_gradient_house"])
    assert FlowValidate().check_strategy(st) == (True, "Pass")
    timing, amount, drops = Strategy.flow_strategy_analysis(st, "example_training_0")
    assert sum(amount) == 1024 and len(amount) == 8 and max(amount) == amount[3] == amount[4]      # bell curve, conserved total


def test_cpu_example_runs():
    r = subprocess.run([sys.executable, "-m", "olearning_sim_b200", "run", os.path.join(ROOT, "examples", "task_mlp_cpu.json"),
                        "--timeout", "200"], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr[-600:]
    assert json.loads(r.stdout.strip().splitlines()[-1])["status"] == "SUCCEEDED"
