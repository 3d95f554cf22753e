"""Drop-in for update_ema_variables (utils/parameters.py:4-8; the (model, ema_model, alpha,
global_step) variant of utils/udaap/utils_mt.py:34-39 is accepted too).  One kernel launch for
all parameter tensors of the model pair instead of two per tensor."""
import weakref

from . import ops

_plans = weakref.WeakKeyDictionary()


def _plan(model, ema_model):
    per_model = _plans.setdefault(ema_model, {})
    plan = per_model.get(id(model))
    if plan is None:
        plan = ops.EmaPlan([p.data for p in model.parameters()], [p.data for p in ema_model.parameters()])
        per_model[id(model)] = plan
    return plan


def update_ema_variables(model, ema_model, args, global_step=None):
    if global_step is not None:                       # utils_mt.py:34-39 signature: (model, ema, alpha, step)
        alpha = min(1 - 1 / (global_step + 1), args)
    else:                                             # parameters.py:6: keyed on the EPOCH
        alpha = min(1 - 1 / (args.epo + 1), args.ema_decay)
    plan = _plan(model, ema_model)
    # parameters may have been re-allocated (optimizer swaps, .to()): refresh the tensor lists
    plan.params = [p.data for p in model.parameters()]
    plan.ema_params = [p.data for p in ema_model.parameters()]
    plan.step(alpha)
