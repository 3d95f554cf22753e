"""Drop-in for the two-teacher assessment / selection of utils/business.py (BusinessUtils):

    assess_pseudo_unc    utils/business.py:16-35     assess_pseudo_unc2   utils/business.py:109-161
    filter_pseudo2       utils/business.py:173-217   preds_mean           utils/business.py:297-300
    filter_pseudo        utils/business.py:49-91

Same signatures and the same list-of-dict records (keys kpID, imageID, kIdx, coord, coord_gt,
coord_legal, error, acc_flag, coord_w1, coord_w2, intDist1, intDist2, extDist, reliability,
enable).  All arithmetic (pairwise dispersions, ensemble weights, errors / PCK flags,
min-max normalisation, the exact global order statistic and the masks) runs in the float64
kernels of K2; the host only assembles the python records the API has to return.
`pseudo_cal_unc` and `pseudo_filter_mixUnc(2)` (the stateful 3-epoch LMA variant, business.py:220-294)
are not covered yet and stay with the reference.
"""
import copy

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
        # same kernels as filter_pseudo2: identical arithmetic as long as no distance equals the 999 sentinel
        s = ops.select_quantile(dist, legal, args.kpsCount, args.reliableThr, args.reliablePCT, args.reliableDistMin)
        dl, rel, en = dist.cpu().tolist(), s["reliability"].cpu().tolist(), s["enable"].cpu().tolist()
        thr = float(s["thr"].item())
        for i, p in enumerate(pseudoArray):
            p["dist"] = dl[i]
            p["coord_legal"] = predsArray_mds1[i]["coord_legal"] and predsArray_mds2[i]["coord_legal"]
            p["reliability"] = rel[i]
        return cls._collect(pseudoArray, rel, en, thr, args)

    @classmethod
    def _collect(cls, pseudoArray, rel, en, thr, args):
        """The selection loop shared by filter_pseudo / filter_pseudo2 (business.py:69-91,195-217)."""
        order = sorted(range(len(pseudoArray)), key=lambda i: rel[i], reverse=True)   # stable, like sorted(..., reverse=True)
        n = args.kpsCount + 1
        selArray, selCounts, selErrs, selAccs = [], [0] * n, [0] * n, [0] * n
        for i in order:
            item = copy.deepcopy(pseudoArray[i])
            if en[i]:
                kID = int(item["kpID"].split("_")[-1])
                item["enable"] = 1
                selCounts[-1] += 1
                selCounts[kID] += 1
                selErrs[-1] += item["error"]
                selErrs[kID] += item["error"]
                selAccs[-1] += item["acc_flag"]
                selAccs[kID] += item["acc_flag"]
            else:
                item["enable"] = 0
            selArray.append(item)
        for idx in range(n):
            if selCounts[idx] > 0:
                selErrs[idx] = selErrs[idx] / selCounts[idx]
                selAccs[idx] = selAccs[idx] / selCounts[idx]
        return selArray, selCounts, selErrs, selAccs, thr

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
