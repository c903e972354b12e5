"""``hvd.SyncBatchNorm`` - batch normalisation with statistics reduced over all workers.

Part of the ``horovod.torch`` surface a user of HorovodRunner expects (SURVEY.md §7 step 3 API
list).  Forward reduces ``[sum(x), sum(x^2), count]`` with one differentiable ``hvd.allreduce``;
autograd then yields the matching allreduce of the gradient statistics in backward.
"""
import torch
from torch.nn.modules.batchnorm import _BatchNorm


class SyncBatchNorm(_BatchNorm):
    def _check_input_dim(self, input):  # pylint: disable=redefined-builtin
        if input.dim() < 2:
            raise ValueError("expected at least 2D input (got %dD input)" % input.dim())

    def forward(self, input):  # pylint: disable=redefined-builtin
        from spark_deep_learning_b200.parallel import hvd
        self._check_input_dim(input)
        if not self.training or not hvd.is_initialized() or hvd.size() == 1:
            return torch.nn.functional.batch_norm(
                input, self.running_mean if not self.training or self.track_running_stats else None,
                self.running_var if not self.training or self.track_running_stats else None,
                self.weight, self.bias, self.training, self._momentum(), self.eps)
        c = input.shape[1]
        dims = [0] + list(range(2, input.dim()))
        x = input.float() if input.dtype in (torch.float16, torch.bfloat16) else input   # fp64 stays fp64
        n_local = x.numel() // c
        stats = torch.cat([x.sum(dims), (x * x).sum(dims),
                           torch.full((1,), float(n_local), device=x.device, dtype=x.dtype)])
        stats = hvd.allreduce(stats, op=hvd.Sum, name=None)
        n = stats[2 * c]
        mean = stats[:c] / n
        var = (stats[c:2 * c] / n - mean * mean).clamp_min(0.0)
        if self.track_running_stats:
            with torch.no_grad():
                m = self._momentum()
                self.running_mean.mul_(1 - m).add_(mean.detach().to(self.running_mean.dtype), alpha=m)
                unbiased = var.detach() * (n / (n - 1).clamp_min(1.0))
                self.running_var.mul_(1 - m).add_(unbiased.to(self.running_var.dtype), alpha=m)
                self.num_batches_tracked += 1
        shape = [1, c] + [1] * (input.dim() - 2)
        y = (x - mean.view(shape)) * torch.rsqrt(var.view(shape) + self.eps)
        if self.affine:
            y = y * self.weight.view(shape).to(x.dtype) + self.bias.view(shape).to(x.dtype)
        return y.to(input.dtype)

    def _momentum(self):
        if self.momentum is None:
            return 1.0 / float(self.num_batches_tracked + 1) if self.num_batches_tracked is not None else 0.0
        return self.momentum
