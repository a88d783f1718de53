"""``ModelManager`` — drop-in for ``footprints/model_manager.py:14-56`` of the reference.

Same constructor arguments, attributes (``model``, ``optimiser``, ``scheduler``) and methods
(``load_model``, ``save_model``) and the same on-disk format (``model.pth`` / ``optimiser.pth``
holding plain ``state_dict``s), so checkpoints are interchangeable with the reference.

Differences, all additive: ``pretrained`` / ``blocks`` / ``compute_dtype`` keyword arguments
(the reference hard-codes ``pretrained=True``, which needs network access); ``load_model`` passes
``map_location`` to ``torch.load`` — the reference passes it to ``load_state_dict`` (model_manager.py:37),
which raises ``TypeError`` on CPU-only hosts.
"""
import os

import torch

from .network import FootprintNetwork


class ModelManager:

    def __init__(self, save_folder=None, use_cuda=True, is_inference=False,
                 learning_rate=1e-4, lr_step_size=10, pretrained=True, blocks=(3, 4, 6, 3),
                 compute_dtype='fp16'):

        self.save_folder = save_folder
        self.use_cuda = use_cuda

        self.model = FootprintNetwork(pretrained=pretrained, blocks=blocks, compute_dtype=compute_dtype)
        if self.use_cuda:
            self.model.cuda()

        if not is_inference:
            # footprints/model_manager.py:27-28: Adam(lr), StepLR(step_size)
            self.optimiser = torch.optim.Adam(self.model.parameters(), lr=learning_rate)
            self.scheduler = torch.optim.lr_scheduler.StepLR(self.optimiser, step_size=lr_step_size)

    def load_model(self, weights_path, load_optimiser=False):

        print('loading model weights from {}...'.format(weights_path))
        map_location = None if torch.cuda.is_available() else 'cpu'
        weights = torch.load(os.path.join(weights_path, 'model.pth'), map_location=map_location)
        self.model.load_state_dict(weights)
        print('successfully loaded weights!')

        if load_optimiser:
            print('loading optimiser...')
            weights = torch.load(os.path.join(weights_path, 'optimiser.pth'), map_location=map_location)
            self.optimiser.load_state_dict(weights)
            print('successfully loaded optimiser!')

    def save_model(self, folder_name):

        save_path = os.path.join(self.save_folder, folder_name)
        print('saving weights to {}...'.format(save_path))
        os.makedirs(save_path, exist_ok=True)

        torch.save(self.model.state_dict(), os.path.join(save_path, 'model.pth'))
        torch.save(self.optimiser.state_dict(), os.path.join(save_path, 'optimiser.pth'))
        print('success!')
