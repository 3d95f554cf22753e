"""Makes the reference's drivers run on the B200 kernels WITHOUT editing or copying them: call
`ubpl_b200.install.install()` after putting the reference checkout on sys.path and BEFORE importing
`projects.*` (the drivers bind names at import time, projects/MT_UBPL.py:20-24).  It replaces
attributes of the reference's own modules:

    utils.losses.{JointMSELoss, JointDistLoss, JointPseudoLoss3, JointDistLoss_mt2}
    utils.augment.AugmentUtils.{affine_back2, affine_back2_classification, fliplr_back_tensor, affine_kps}
    utils.process.ProcessUtils.{kps_fromHeatmap, kps_fromHeatmap_mul, kps_fromHeatmap2, kps_heatmap,
                                kps_heatmap_mulKps, kps_getLabeledCount, features_cov}
    utils.evaluation.EvaluationUtils.{uncertainty_fromDistance, acc_pck}
    utils.business.BusinessUtils.{assess_pseudo_unc, assess_pseudo_unc2, filter_pseudo, filter_pseudo2,
                                  pseudo_cal_unc, pseudo_filter_mixUnc, pseudo_filter_mixUnc2, preds_mean}
    utils.parameters.update_ema_variables   (and utils.udaap.utils_mt.update_ema_variables)
"""
import importlib

from . import augment, business, evaluation, losses, parameters, process

PATCHES = {
    "utils.losses": {n: getattr(losses, n) for n in ("JointMSELoss", "JointDistLoss", "JointPseudoLoss3", "JointDistLoss_mt2")},
    "utils.parameters": {"update_ema_variables": parameters.update_ema_variables},
}
CLASS_PATCHES = {
    ("utils.augment", "AugmentUtils"): (augment.AugmentUtils, ("affine_back2", "affine_back2_classification", "fliplr_back_tensor",
                                                               "affine_kps")),
    ("utils.process", "ProcessUtils"): (process.ProcessUtils, ("kps_fromHeatmap", "kps_fromHeatmap_mul", "kps_fromHeatmap2",
                                                               "kps_heatmap", "kps_heatmap_mulKps", "kps_getLabeledCount",
                                                               "features_cov")),
    ("utils.evaluation", "EvaluationUtils"): (evaluation.EvaluationUtils, ("uncertainty_fromDistance", "acc_pck")),
    ("utils.business", "BusinessUtils"): (business.BusinessUtils, ("assess_pseudo_unc", "assess_pseudo_unc2", "filter_pseudo", "filter_pseudo2",
                                                                    "pseudo_cal_unc", "pseudo_filter_mixUnc", "pseudo_filter_mixUnc2",
                                                                    "preds_mean")),
}


_saved = []          # (object, attribute name, original value or _MISSING) of the last install()
_MISSING = object()


def _swap(obj, name, value):
    _saved.append((obj, name, obj.__dict__.get(name, _MISSING)))
    setattr(obj, name, value)


def install():
    """Patch the already-importable reference modules in place; returns the list of patched names."""
    done = []
    for mod_name, attrs in PATCHES.items():
        mod = importlib.import_module(mod_name)
        for k, v in attrs.items():
            _swap(mod, k, v)
            done.append("%s.%s" % (mod_name, k))
    for (mod_name, cls_name), (src, names) in CLASS_PATCHES.items():
        cls = getattr(importlib.import_module(mod_name), cls_name)
        for n in names:
            _swap(cls, n, getattr(src, n))            # bound to THIS package's class: its helpers stay reachable
            done.append("%s.%s.%s" % (mod_name, cls_name, n))
    try:
        um = importlib.import_module("utils.udaap.utils_mt")
        _swap(um, "update_ema_variables", parameters.update_ema_variables)
        done.append("utils.udaap.utils_mt.update_ema_variables")
    except Exception:
        pass
    return done


def uninstall():
    """Puts back what install() replaced (tests compare the reference's own classes with the patched ones in one
    process); returns the number of attributes restored."""
    n = 0
    while _saved:
        obj, name, old = _saved.pop()
        if old is _MISSING:
            try:
                delattr(obj, name)
            except AttributeError:
                pass
        else:
            setattr(obj, name, old)
        n += 1
    return n
