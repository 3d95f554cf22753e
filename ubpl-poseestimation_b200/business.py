"""Drop-in for the two-teacher assessment / selection of utils/business.py (BusinessUtils):

    assess_pseudo_unc    utils/business.py:16-35     assess_pseudo_unc2   utils/business.py:109-161
    filter_pseudo2       utils/business.py:173-217   preds_mean           utils/business.py:297-300
    filter_pseudo        utils/business.py:49-91

Same signatures and the same list-of-dict records (keys kpID, imageID, kIdx, coord, coord_gt,
coord_legal, error, acc_flag, coord_w1, coord_w2, intDist1, intDist2, extDist, reliability,
enable).  All arithmetic (pairwise dispersions, ensemble weights, errors / PCK flags,
min-max normalisation, the exact global order statistic and the masks) runs in the float64
kernels of K2; the host only assembles the python records the API has to return.
`pseudo_cal_unc` / `pseudo_filter_mixUnc(2)` (the stateful 3-epoch LMA variant, business.py:220-294): the
O(B*J*A^2) pairwise dispersions, the teacher-to-teacher distances and the view means run in ubpl_mix_dists; the
LMA cache stays in args.mdsN_lma_cache exactly as the reference keeps it (same list of dicts), and the two
distances whose radicand is not an integer (error against a fractional gt, aExtDist between view means) are
taken with CPython's own `** 0.5` on the host so that every record field is bit-identical.
"""
import copy
import numpy as np
import math
import torch

from . import ops


def _cuda(t, dtype=torch.float32):
    t = torch.as_tensor(t).detach().to(dtype)
    return t if t.is_cuda else t.cuda()


class BusinessUtils:
    @classmethod
    def preds_mean(cls, preds1, preds2):
        """utils/business.py:297-300 (tiny [B,J,2] tensors: plain torch)."""
        return torch.mean(torch.stack([preds1, preds2], dim=-1), dim=-1)

    @classmethod
    def _records(cls, imageIDs, preds, gt, err, acc):
        """preds [P,B,J,2] float32 (python lists), gt [B,J,G] -> P lists of B*J dict records."""
        P, B, J = len(preds), len(preds[0]), len(preds[0][0])
        out = []
        for p in range(P):
            rows = []
            for b in range(B):
                imageID = imageIDs[b]
                for j in range(J):
                    c = preds[p][b][j]
                    rows.append({"kpID": "{}_{}".format(imageID, j), "imageID": imageID, "kIdx": j, "coord": c,
                                 "coord_gt": gt[b][j], "coord_legal": 1.0 if c[0] >= 0 and c[1] >= 0 else 0.,
                                 "error": err[p][b][j], "acc_flag": acc[p][b][j]})
            out.append(rows)
        return out

    @classmethod
    def assess_pseudo_unc(cls, imageIDs, test_kpsMap, preds, args):
        """utils/business.py:16-35: per prediction set, per key point: coord, gt, legality, error, PCK flag."""
        gt = _cuda(test_kpsMap)
        pr = torch.stack([_cuda(p) for p in preds])                       # [P,B,J,2]
        err, acc = ops.coord_error(pr, gt, args.pck_ref, args.pck_thr)
        return cls._records(imageIDs, pr.double().cpu().tolist(), test_kpsMap.cpu().data.numpy().tolist(),
                            err.cpu().tolist(), acc.cpu().tolist())

    @classmethod
    def assess_pseudo_unc2(cls, imageIDs, test_kpsMap, ori_predsArray, augs_predsArray, args):
        """utils/business.py:109-161."""
        ori_assess = cls.assess_pseudo_unc(imageIDs, test_kpsMap, ori_predsArray, args)
        augs_assess = [cls.assess_pseudo_unc(imageIDs, test_kpsMap, a, args) for a in augs_predsArray]
        K = args.br_inferAugNum
        a1 = torch.stack([_cuda(p) for p in augs_predsArray[0]])
        a2 = torch.stack([_cuda(p) for p in augs_predsArray[1]])
        if a1.shape[0] != K or a2.shape[0] != K:
            # business.py:150 indexes the first br_inferAugNum views for extDist but averages intDist over all
            raise ValueError("len(augs_predsArray[m]) must equal args.br_inferAugNum")
        ad = ops.assess_dual(_cuda(ori_predsArray[0]), _cuda(ori_predsArray[1]), _cuda(ori_predsArray[-1]), a1, a2)
        gt = _cuda(test_kpsMap)
        err, acc = ops.coord_error(ad["coord"].unsqueeze(0), gt, args.pck_ref, args.pck_thr)
        host = {k: ad[k].cpu().tolist() for k in ("legal", "intDist1", "intDist2", "extDist", "w1", "w2", "coord")}
        if int(ad["zero_div"].item()) > 0:
            raise ZeroDivisionError("float division by zero")            # business.py:135 when both intDists are 0
        err, acc = err[0].cpu().tolist(), acc[0].cpu().tolist()
        B, J = len(host["legal"]), len(host["legal"][0])
        pseudoArray = []
        for i, base in enumerate(ori_assess[-1]):
            b, j = divmod(i, J)
            it = copy.deepcopy(base)
            ok = host["intDist1"][b][j] != 999 or host["extDist"][b][j] != 999
            it["coord_w1"], it["coord_w2"] = host["w1"][b][j], host["w2"][b][j]
            # the reference stores the int sentinel 999 when a group is illegal (business.py:123)
            for k_rec, k_dev in (("intDist1", "intDist1"), ("intDist2", "intDist2"), ("extDist", "extDist")):
                v = host[k_dev][b][j]
                it[k_rec] = v if ok else 999
            it["coord_legal"] = host["legal"][b][j]
            if ok:
                it["coord"] = host["coord"][b][j]
            it["error"], it["acc_flag"] = err[b][j], acc[b][j]
            pseudoArray.append(it)
        return pseudoArray, ori_assess, augs_assess

    @classmethod
    def filter_pseudo(cls, predsArraies, args):
        """utils/business.py:49-91: reliability from the distance between the two teachers' coordinates
        (np.min / np.max normalisation, dist_min floored by reliableDistMin), global quantile, strict >."""
        predsArray_mds1, predsArray_mds2, pseudoArray = predsArraies
        if len(pseudoArray) == 0:
            raise ValueError("zero-size array to reduction operation maximum which has no identity")   # np.max([])
        c1 = torch.tensor([p["coord"] for p in predsArray_mds1], dtype=torch.float64).cuda()
        c2 = torch.tensor([p["coord"] for p in predsArray_mds2], dtype=torch.float64).cuda()
        legal = torch.tensor([1.0 if (a["coord_legal"] and b["coord_legal"]) else 0.0
                              for a, b in zip(predsArray_mds1, predsArray_mds2)], dtype=torch.float64).cuda()
        dist = ops.pair_distance(c1, c2)
        dl = dist.cpu().tolist()
        # The selector kernel is filter_pseudo2's: it treats 999 as "no distance" and replaces an all-zero maximum by
        # 999 (business.py:176-182).  filter_pseudo normalises with the plain np.max / np.min (:59-60), so the two agree
        # unless a distance reaches 999 (impossible between coordinates of a 256-pixel frame) or all distances are
        # equal, where the reference divides by zero (:63) -- both are refused here instead of being answered differently.
        if max(dl) >= 999.0:
            raise ValueError("filter_pseudo: a teacher-to-teacher distance of %g pixels (>= 999, the sentinel of "
                             "filter_pseudo2's kernel) cannot be ranked like utils/business.py:59-63 does" % max(dl))
        if max(dl) == min(min(dl), args.reliableDistMin) and any(bool(v) for v in legal.cpu().tolist()):
            raise ZeroDivisionError("float division by zero")            # (dist-dist_min)/(dist_max-dist_min), business.py:63
        s = ops.select_quantile(dist, legal, args.kpsCount, args.reliableThr, args.reliablePCT, args.reliableDistMin)
        rel, en = s["reliability"].cpu().tolist(), s["enable"].cpu().tolist()
        thr = float(s["thr"].item())
        for i, p in enumerate(pseudoArray):
            p["dist"] = dl[i]
            p["coord_legal"] = predsArray_mds1[i]["coord_legal"] and predsArray_mds2[i]["coord_legal"]
            p["reliability"] = rel[i]
        return cls._collect(pseudoArray, rel, en, thr, args)

    @staticmethod
    def _tally(items, enabled, kps_count):
        """Per-joint and overall (count, mean error, mean PCK flag) of the enabled records, slot kps_count = all joints:
        what the reference accumulates record by record (utils/business.py:69-91,195-217,243-261).  np.bincount adds
        its weights in array order, i.e. in the order of `items`, so the float sums round like the reference's running
        `+=`; joints without a selected record keep the integer 0 the reference initialises them with."""
        n = kps_count + 1
        on = np.flatnonzero(np.asarray(enabled, dtype=bool))
        joint = np.array([int(items[i]["kpID"].rsplit("_", 1)[-1]) for i in on], dtype=np.intp)
        whole = np.full(len(on), kps_count, dtype=np.intp)
        counts = (np.bincount(joint, minlength=n)[:n] + np.bincount(whole, minlength=n)[:n]).tolist()
        means = []
        for field in ("error", "acc_flag"):
            w = np.array([items[i][field] for i in on], dtype=np.float64)
            tot = np.bincount(joint, weights=w, minlength=n)[:n] + np.bincount(whole, weights=w, minlength=n)[:n]
            means.append([float(tot[k]) / counts[k] if counts[k] > 0 else 0 for k in range(n)])
        return counts, means[0], means[1]

    @classmethod
    def _emit(cls, items, enabled, kps_count):
        """Deep copies of `items` (the reference hands out copies) carrying their `enable` flag, plus the tallies."""
        out = []
        for it, e in zip(items, enabled):
            rec = copy.deepcopy(it)
            rec["enable"] = 1 if e else 0
            out.append(rec)
        return (out,) + cls._tally(items, enabled, kps_count)

    @classmethod
    def _collect(cls, pseudoArray, rel, en, thr, args):
        """Output of filter_pseudo / filter_pseudo2: the records ordered by falling reliability (stable, like the
        reference's sorted(..., reverse=True)), their enable flags from the device masks, the tallies, the threshold."""
        order = sorted(range(len(pseudoArray)), key=rel.__getitem__, reverse=True)
        return cls._emit([pseudoArray[i] for i in order], [en[i] for i in order], args.kpsCount) + (thr,)

    @classmethod
    def filter_pseudo2(cls, pseudoArray, args):
        """utils/business.py:173-217: min/max-normalised reliability, global quantile threshold, strict >."""
        if len(pseudoArray) == 0:
            raise IndexError("list index out of range")                  # business.py:45 on an empty list
        ext = torch.tensor([float(p["extDist"]) for p in pseudoArray], dtype=torch.float64).cuda()
        legal = torch.tensor([float(p["coord_legal"]) for p in pseudoArray], dtype=torch.float64).cuda()
        s = ops.select_quantile(ext, legal, args.kpsCount, args.reliableThr, args.reliablePCT, args.reliableDistMin)
        rel = s["reliability"].cpu().tolist()
        en = s["enable"].cpu().tolist()
        thr = float(s["thr"].item())
        for p, r in zip(pseudoArray, rel):
            p["reliability"] = r                                         # the reference mutates its input too (:189)
        return cls._collect(pseudoArray, rel, en, thr, args)


    # ---- a13: pseudo_cal_unc / pseudo_filter_mixUnc(2)  (utils/business.py:220-294, 302-406) ------------------
    @staticmethod
    def _pydist(c1, c2):
        return ((c1[0] - c2[0]) ** 2 + (c1[1] - c2[1]) ** 2) ** 0.5      # utils/process.py:53-54, CPython pow

    _LMA = (0.5, 0.3, 0.2)                                                # weights of the newest, previous, oldest epoch

    @classmethod
    def _lma_variables(cls, sources):
        """utils/business.py:395-405: moving average over the last (up to) three epochs, newest weighted most; with
        two epochs the newest takes the first two weights; 999 before the first.  Same products and left-to-right
        sums as the reference, so the cached values are bit-identical."""
        recent = sources[:-4:-1]                                          # newest first
        if not recent:
            return 999.0
        if len(recent) == 1:
            return recent[0]
        w = cls._LMA if len(recent) == 3 else (cls._LMA[0] + cls._LMA[1], cls._LMA[2])
        acc = recent[0] * w[0]
        for x, wk in zip(recent[1:], w[1:]):
            acc = acc + x * wk
        return acc

    @staticmethod
    def _calUncValue(mixDist):
        return 1.0 - math.exp(-mixDist / 5)                              # utils/business.py:375-376

    @classmethod
    def pseudo_cal_unc(cls, imageIDs, preds_gt, preds_mds1, scores_mds1, augPredsArray_mds1, augScoresArray_mds1,
                       preds_mds2, scores_mds2, augPredsArray_mds2, augScoresArray_mds2, args):
        """utils/business.py:220-234: per key point and teacher the record of _initKSample (:302-318) completed by
        _calKSampleExterData (:320-346); args.mds1_lma_cache / args.mds2_lma_cache are appended to like the
        reference does (kpID-keyed dicts), here through a kpID index instead of a linear search per key point."""
        a1, a2 = _cuda(augPredsArray_mds1), _cuda(augPredsArray_mds2)
        B, J = a1.shape[0], a1.shape[1]
        d = ops.mix_dists(_cuda(preds_mds1), _cuda(scores_mds1), a1, _cuda(preds_mds2), _cuda(scores_mds2), a2)
        host = {k: v.cpu().tolist() for k, v in d.items() if v is not None}
        gt = preds_gt.cpu().data.numpy().tolist()
        norms = [cls._pydist(gt[b][args.pck_ref[0]], gt[b][args.pck_ref[1]]) for b in range(B)]
        out = []
        for m, (preds, scores, aug, aug_s) in enumerate(((preds_mds1, scores_mds1, augPredsArray_mds1, augScoresArray_mds1),
                                                         (preds_mds2, scores_mds2, augPredsArray_mds2, augScoresArray_mds2))):
            pl = preds.cpu().data.numpy().tolist()
            al = aug.cpu().data.numpy().tolist()
            s0 = scores[0].cpu().data.numpy().tolist()                  # batch row 0 for every sample (:310-311)
            as0 = aug_s[0].cpu().data.numpy().tolist()
            tag = str(m + 1)
            recs = []
            for b in range(B):
                for j in range(J):
                    err = cls._pydist(pl[b][j], gt[b][j])
                    k_scores = [max(0.0, min(1.0, x)) for x in as0[j]] + [max(0.0, min(1.0, s0[j]))]
                    recs.append({"kpID": "{}_{}".format(imageIDs[b], j), "coord": pl[b][j], "coord_gt": gt[b][j], "error": err,
                                 "acc_flag": 1 if err / norms[b] < args.pck_thr else 0, "coords_aug": al[b][j],
                                 "coord_aug": host["caug" + tag][b][j], "scores": k_scores, "score": k_scores[-1],
                                 "intDist": host["int" + tag][b][j]})
            out.append(recs)
        caches = (args.mds1_lma_cache, args.mds2_lma_cache)
        index = [{it["kpID"]: it for it in c} for c in caches]
        for i in range(B * J):
            b, j = divmod(i, J)
            s1, s2 = out[0][i], out[1][i]
            s1["extDist"] = s2["extDist"] = host["ext"][b][j]
            s1["aExtDist"] = s2["aExtDist"] = cls._pydist(s1["coord_aug"], s2["coord_aug"])
            for smp, cache, idx in ((s1, caches[0], index[0]), (s2, caches[1], index[1])):
                tgt = idx.get(smp["kpID"])
                if tgt is None:                                          # getLMAfromCache (:348-355)
                    tgt = {"kpID": smp["kpID"], "intDist": [], "extDist": [], "aExtDist": [], "intDist_lma": [],
                           "extDist_lma": [], "aExtDist_lma": []}
                    cache.append(tgt)
                    idx[smp["kpID"]] = tgt
                for q in ("intDist", "extDist", "aExtDist"):
                    tgt[q].append(smp[q])
                for q in ("intDist", "extDist", "aExtDist"):
                    smp[q + "_lma"] = cls._lma_variables(tgt[q])
                    tgt[q + "_lma"].append(smp[q + "_lma"])
                smp["mixDist"] = smp["intDist_lma"] + ((smp["extDist_lma"] + smp["aExtDist_lma"]) / 2
                                                       if smp["extDist_lma"] > 0 else smp["aExtDist_lma"])
            for smp in (s1, s2):
                for q in ("intDist", "extDist", "aExtDist"):
                    smp[q + "OK"] = 1 if smp[q] <= args.distThrMax else 0
                    smp[q + "OK_lma"] = 1 if smp[q + "_lma"] <= args.distThrMax else 0
                ok = smp["intDistOK_lma"] > 0 and smp["extDistOK_lma"] > 0 and smp["aExtDistOK_lma"] > 0
                smp["unc"] = cls._calUncValue(smp["mixDist"]) if ok else 999.0
        return out[0], out[1]

    @classmethod
    def _mix_collect(cls, pseudoArray, uncThr, args):
        """Output of pseudo_filter_mixUnc / pseudo_filter_mixUnc2 (business.py:243-261, 271-293): records in their given
        order, enabled where unc <= uncThr."""
        return cls._emit(pseudoArray, [it["unc"] <= uncThr for it in pseudoArray], args.kpsCount)

    @classmethod
    def pseudo_filter_mixUnc(cls, pseudoArray, args):
        """utils/business.py:237-261: enable = unc <= 1-exp(-3*distThrMax/5)."""
        uncThr = cls._calUncValue(args.distThrMax * 3)
        return cls._mix_collect(pseudoArray, uncThr, args) + (uncThr,)

    @classmethod
    def pseudo_filter_mixUnc2(cls, pseudoArray, args):
        """utils/business.py:264-294: items scoring below the median score (:357-364) get unc = 999 first; the
        reference mutates its input records (scoreOK, unc), so does this."""
        ranked = sorted((it["score"] for it in pseudoArray), reverse=True)
        scoreThr = ranked[int((len(ranked) - 1) * 0.5)]                   # the median score (:357-364)
        for it in pseudoArray:
            low = it["score"] < scoreThr
            it["scoreOK"] = 0 if low else 1
            if low:
                it["unc"] = 999.0
        uncThr = cls._calUncValue(args.distThrMax * 3)
        return cls._mix_collect(pseudoArray, uncThr, args) + (scoreThr, uncThr)
