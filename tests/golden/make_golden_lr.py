#!/usr/bin/env python
"""Golden (step -> learning rate) points from the reference's OWN CosineWarmup schedule driven the way
its LearningRateScheduler callback drives it (util/learning_rate_schedule.py:55-79,189-277) and its
KLRegularizerSchedule (util/trainer.py:298-316), executed on a minimal stand-in for the few
TensorFlow names that file touches (float arithmetic only).  Build container only:

    python tests/golden/make_golden_lr.py        ->  tests/golden/ref_lr_schedule.json
"""
import importlib.util
import json
import math
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/util/learning_rate_schedule.py"


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Scope:
    def __init__(self, *a, **k): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False


class _F(float):
    dtype = "float"


def install():
    ops = _mod("tensorflow.python.framework.ops", name_scope_v2=_Scope, Tensor=_F,
               convert_to_tensor_v2=lambda v, name=None: _F(v))
    math_ops = _mod("tensorflow.python.ops.math_ops", cast=lambda v, dtype: float(v))
    _mod("tensorflow.python.ops", math_ops=math_ops)
    _mod("tensorflow.python.framework", ops=ops)
    _mod("tensorflow.python.util.tf_export", keras_export=lambda *a, **k: (lambda cls: cls))
    _mod("tensorflow.python.util")
    backend = _mod("tensorflow.python.keras.backend", eval=float, get_value=float, set_value=lambda a, b: None)
    _mod("tensorflow.python.keras", backend=backend)
    _mod("tensorflow.python")
    schedules = types.SimpleNamespace(LearningRateSchedule=object)
    keras = types.SimpleNamespace(callbacks=types.SimpleNamespace(Callback=object), backend=backend,
                                  optimizers=types.SimpleNamespace(schedules=schedules))
    _mod("tensorflow", keras=keras, cos=lambda v: _F(math.cos(v)), constant=_F,
         multiply=lambda a, b: _F(a * b))


def main():
    install()
    spec = importlib.util.spec_from_file_location("ref_lr_schedule", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cases = []
    # (train_size, num_gpus, batch, epochs, warmup_epochs): README CIFAR-10 run, an 8-GPU run, a toy run
    for train_size, gpus, batch, epochs, warm in [(50000, 1, 40, 400, 5), (50000, 8, 40, 400, 5), (320, 2, 16, 12, 2)]:
        spe = int(train_size / (gpus * batch))                                  # tr.py:139
        decay = epochs - warm                                                   # tr.py:140
        sched = ref.CosineWarmup(initial_learning_rate=1e-2, min_learning_rate=1e-4, warmup_epochs=warm,
                                 decay_epochs=decay, steps_per_epoch=spe, verbose=False)
        total = epochs * spe
        probe = sorted({1, 2, spe, warm * spe - 1, warm * spe, warm * spe + 1, (warm + 1) * spe - 1,
                        (warm + 1) * spe, (warm + decay // 2) * spe, (warm + decay) * spe - 1,
                        (warm + decay) * spe, total, total + spe, total + 3 * spe} | set(range(1, total, max(1, total // 23))))
        pts = []
        step = 0
        want = set(probe)
        for s in range(1, max(probe) + 1):                                      # the callback's global_step (1-based)
            step = s
            lr = float(sched(step))
            if s in want:
                pts.append([s, lr])
        cases.append({"train_size": train_size, "num_gpus": gpus, "batch_size": batch, "epochs": epochs,
                      "lr_warmup_epochs": warm, "learning_rate": 1e-2, "lr_min_learning_rate": 1e-4,
                      "steps_per_epoch": spe, "lr_decay_epochs": decay, "points": pts})
    with open(os.path.join(HERE, "ref_lr_schedule.json"), "w") as f:
        json.dump({"source": "reference util/learning_rate_schedule.py CosineWarmup via LearningRateScheduler "
                             "(global_step incremented before the call)", "cases": cases}, f)
    print("wrote ref_lr_schedule.json:", [len(c["points"]) for c in cases])


if __name__ == "__main__":
    main()
