"""One K-Planes training iteration, mirroring ``Trainer.train_iteration`` (NS/engine/trainer.py:382-412) and the
callback calls around it (:212, :221): callbacks BEFORE -> zero_grad -> forward (collider + get_outputs) ->
metrics -> loss dict -> backward -> [gradient all-reduce] -> Adam -> scheduler -> callbacks AFTER."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from ..cameras.rays import RayBundle
from ..distributed import GradBucket, world_info
from ..models.kplanes import KPlanesModel, TrainingCallbackLocation
from .optimizers import Optimizers


class TrainStep:
    def __init__(self, model: KPlanesModel, max_steps: int = 30000, lr: float = 1e-2, eps: float = 1e-12,
                 warm_up_end: int = 512, data_parallel: bool = False) -> None:
        self.model = model
        self.optimizers = Optimizers(model.get_param_groups(), lr=lr, eps=eps, warm_up_end=warm_up_end, max_steps=max_steps)
        self.callbacks = model.get_training_callbacks(None)
        self.step = 0
        self.rank, self.world = world_info()
        self.bucket: Optional[GradBucket] = None
        if data_parallel and self.world > 1:
            self.bucket = GradBucket([p for ps in model.get_param_groups().values() for p in ps])

    def __call__(self, ray_bundle: RayBundle, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        model = self.model
        model.train()
        for cb in self.callbacks:
            cb.run_callback_at_location(self.step, TrainingCallbackLocation.BEFORE_TRAIN_ITERATION)
        if self.bucket is not None:
            self.bucket.attach_zeroed()
        else:
            self.optimizers.zero_grad_all()
        outputs = model(ray_bundle)
        metrics = model.get_metrics_dict(outputs, batch)
        loss_dict = model.get_loss_dict(outputs, batch, metrics)
        loss = sum(loss_dict.values())
        loss.backward()
        grad_scale = 1.0
        if self.bucket is not None:
            self.bucket.all_reduce()
            grad_scale = 1.0 / self.world
        self.optimizers.optimizer_step_all(grad_scale=grad_scale)
        self.optimizers.scheduler_step_all(self.step)
        for cb in self.callbacks:
            cb.run_callback_at_location(self.step, TrainingCallbackLocation.AFTER_TRAIN_ITERATION)
        self.step += 1
        loss_dict["loss"] = loss.detach()
        return loss_dict
