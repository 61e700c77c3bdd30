"""EoT wrapper -- same behaviour as /root/reference/src/defenses/wrappers.py:4-24 (repeat, forward, mean over
the replicas), plus a batched variant for B > 1 images that keeps every image's replicas on
one GPU and averages per image."""
import torch


class EoTWrapper(torch.nn.Module):
    def __init__(self, model: torch.nn.Module, eot_steps: int):
        super().__init__()
        self.model = model
        self.eot_steps = eot_steps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: (1, 3, h, w) -> (1, n_classes)   [reference semantics]
           x: (B, 3, h, w) -> (B, n_classes)   [batched extension: replicas are interleaved per image]"""
        b = x.shape[0]
        if b == 1:
            x = x.repeat(self.eot_steps, 1, 1, 1)
            preds = self.model(x)
            return torch.mean(preds, dim=0, keepdim=True)
        xr = x.repeat_interleave(self.eot_steps, dim=0)
        preds = self.model(xr)
        return preds.view(b, self.eot_steps, -1).mean(dim=1)
