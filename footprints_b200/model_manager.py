"""``ModelManager`` -- drop-in for ``footprints/model_manager.py:14-56`` of the reference.

Constructor arguments, attributes (``model``, ``optimiser``, ``scheduler``), the two methods and the on-disk
layout (``<folder>/model.pth`` and ``<folder>/optimiser.pth``, plain ``state_dict``s) are the reference's, so that
``footprints/training/train.py`` and ``footprints/predict_simple.py`` run on top of it unmodified and checkpoints
move freely between the two code bases.

Additions (keyword-only in spirit; the reference's positional order is untouched):
``pretrained`` / ``blocks`` / ``compute_dtype`` (the reference hard-codes ``pretrained=True``, i.e. a download);
``attach_trainer`` -- when the B200 fast path (``footprints_b200.trainer.Trainer``: flat buffers, fused Adam) does
the optimising, the Adam moments live there, and ``save_model`` / ``load_model(load_optimiser=True)`` go through
its ``state_dict`` (torch.optim.Adam format) so that a resumed run keeps its moments and bias correction.
The reference passes ``map_location`` to ``load_state_dict`` (model_manager.py:37, a TypeError on CPU-only hosts);
here it goes to ``torch.load`` where it belongs.
"""
import os

import torch

from .network import FootprintNetwork

_MODEL_FILE, _OPT_FILE = 'model.pth', 'optimiser.pth'


class ModelManager:

    def __init__(self, save_folder=None, use_cuda=True, is_inference=False, learning_rate=1e-4, lr_step_size=10,
                 pretrained=True, blocks=(3, 4, 6, 3), compute_dtype='fp16'):
        self.save_folder, self.use_cuda = save_folder, use_cuda
        self.trainer = None
        self.model = FootprintNetwork(pretrained=pretrained, blocks=blocks, compute_dtype=compute_dtype)
        if use_cuda:
            self.model.cuda()
        if is_inference:
            return
        # Adam(lr) with torch defaults and StepLR(step_size), as model_manager.py:27-28
        self.optimiser = torch.optim.Adam(self.model.parameters(), lr=learning_rate)
        self.scheduler = torch.optim.lr_scheduler.StepLR(self.optimiser, step_size=lr_step_size)

    def attach_trainer(self, trainer):
        """Route optimiser checkpoints through ``trainer`` (see the module docstring)."""
        self.trainer = trainer
        return trainer

    def _read(self, folder, name):
        where = None if torch.cuda.is_available() else 'cpu'
        return torch.load(os.path.join(folder, name), map_location=where)

    def load_model(self, weights_path, load_optimiser=False):
        self.model.load_state_dict(self._read(weights_path, _MODEL_FILE))
        self.model._weights_epoch = getattr(self.model, '_weights_epoch', 0) + 1
        if load_optimiser:
            target = self.trainer if self.trainer is not None else self.optimiser
            target.load_state_dict(self._read(weights_path, _OPT_FILE))

    def save_model(self, folder_name):
        out_dir = os.path.join(self.save_folder, folder_name)
        os.makedirs(out_dir, exist_ok=True)
        source = self.trainer if self.trainer is not None else self.optimiser
        torch.save(self.model.state_dict(), os.path.join(out_dir, _MODEL_FILE))
        torch.save(source.state_dict(), os.path.join(out_dir, _OPT_FILE))
