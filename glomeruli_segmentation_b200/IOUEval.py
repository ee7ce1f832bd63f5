"""`iouEval` with the reference's interface (`module/common/IOUEval.py:7-69`): addBatch / getMetric /
getMetricRight.  CUDA uint8 label maps are histogrammed on the GPU (espnet_confusion_hist, no D2H of
the masks); numpy / CPU inputs go through the same bincount arithmetic the reference uses, on the host
(that is bookkeeping on 25 numbers, not the hot path)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class iouEval:
    def __init__(self, nClasses):
        self.nClasses = nClasses
        self.reset()

    def reset(self):
        self.overall_acc = 0
        self.per_class_acc = np.zeros(self.nClasses, dtype=np.float32)
        self.per_class_iu = np.zeros(self.nClasses, dtype=np.float32)
        self.mIOU = 0
        self.batchCount = 0
        self.hist = np.zeros((self.nClasses, self.nClasses), dtype=np.int64)

    # IOUEval.py:19-21
    def fast_hist(self, a, b):
        n = self.nClasses
        if isinstance(a, torch.Tensor) and a.is_cuda:
            return self._gpu_hist(b, a)
        a = np.asarray(a).reshape(-1)
        b = np.asarray(b).reshape(-1)
        k = (a >= 0) & (a < n)
        return np.bincount(n * a[k].astype(int) + b[k], minlength=n ** 2).reshape(n, n)

    def _gpu_hist(self, predict: torch.Tensor, gth: torch.Tensor) -> np.ndarray:
        n = self.nClasses
        if predict.dtype != torch.uint8 or gth.dtype != torch.uint8:
            raise RuntimeError("GPU iouEval wants uint8 label maps (what the arg-max kernel writes)")
        predict, gth = predict.contiguous(), gth.contiguous()
        hist = torch.zeros(n * n, dtype=torch.int64, device=predict.device)
        st = torch.cuda.current_stream(predict.device).cuda_stream
        _lib.check(_lib.lib().espnet_confusion_hist(predict.data_ptr(), gth.data_ptr(), predict.numel(), n, hist.data_ptr(), st),
                   None, "espnet_confusion_hist")
        return hist.cpu().numpy().reshape(n, n)

    def compute_hist(self, predict, gth):
        return self.fast_hist(gth, predict)

    # IOUEval.py:27-53
    def addBatch(self, predict, gth):
        if isinstance(predict, torch.Tensor) and not predict.is_cuda:
            predict, gth = predict.numpy(), gth.numpy()
        hist = self.compute_hist(predict, gth)
        self.hist = self.hist + hist
        epsilon = 0.00000001
        d = np.diag(hist)
        overall_acc = d.sum() / (hist.sum() + epsilon)
        per_class_acc = d / (hist.sum(1) + epsilon)
        per_class_iu = d / (hist.sum(1) + hist.sum(0) - d + epsilon)
        self.overall_acc += overall_acc
        self.per_class_acc += per_class_acc
        self.per_class_iu += per_class_iu
        self.mIOU += np.nanmean(per_class_iu)
        self.batchCount += 1
        return hist

    # IOUEval.py:55-61
    def getMetric(self):
        return (self.overall_acc / self.batchCount, self.per_class_acc / self.batchCount,
                self.per_class_iu / self.batchCount, self.mIOU / self.batchCount)

    # IOUEval.py:63-69
    def getMetricRight(self):
        epsilon = 0.00000001
        d = np.diag(self.hist)
        overall_acc = d.sum() / (self.hist.sum() + epsilon)
        per_class_acc = d / (self.hist.sum(1) + epsilon)
        per_class_iu = d / (self.hist.sum(1) + self.hist.sum(0) - d + epsilon)
        return overall_acc, per_class_acc, per_class_iu, np.nanmean(per_class_iu)
