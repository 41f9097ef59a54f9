"""ActorLearner: mirror of actor_learner.py:11-127 (learner base: optimizer state, clip mode, checkpoints,
learning-rate schedule, reward clipping).  The RMSProp slots, the clip and the apply step live on the GPU
(RolloutEngine / optim.cu); this class keeps the reference's constructor signature, attribute names and
helper methods."""
import logging
import os
from multiprocessing import Process

import numpy as np
import torch

from .engine import RolloutEngine
from .session import Session, Saver

CHECKPOINT_INTERVAL = 1000000


class ActorLearner(Process):

    def __init__(self, network_creator, environment_creator, args):

        super(ActorLearner, self).__init__()

        self.global_step = 0

        self.max_local_steps = args.max_local_steps
        self.num_actions = args.num_actions
        self.initial_lr = args.initial_lr
        self.lr_annealing_steps = args.lr_annealing_steps
        self.emulator_counts = args.emulator_counts
        self.device = args.device
        self.debugging_folder = args.debugging_folder
        self.network_checkpoint_folder = os.path.join(self.debugging_folder, 'checkpoints/')
        self.optimizer_checkpoint_folder = os.path.join(self.debugging_folder, 'optimizer_checkpoints/')
        self.last_saving_step = 0

        # one process per GPU: rank r owns environments [r*N/G, (r+1)*N/G)  
        self.world_size = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
        self.rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
        if self.emulator_counts % self.world_size != 0:
            raise ValueError('emulator_counts must be divisible by the number of GPUs')
        self.local_emulator_counts = self.emulator_counts // self.world_size
        first = self.rank * self.local_emulator_counts

        self.emulators = np.asarray([environment_creator.create_environment(first + i)
                                     for i in range(self.local_emulator_counts)])
        self.max_global_steps = args.max_global_steps
        self.gamma = args.gamma
        self.game = args.game
        self.network = network_creator()

        # RMSProp (decay=alpha, epsilon=e, momentum 0), global-norm clip: actor_learner.py:31-70
        self.engine = RolloutEngine(self.network, self.local_emulator_counts, self.max_local_steps, gamma=args.gamma,
                                    rho=args.alpha, eps=args.e, momentum=0.0, clip_norm=args.clip_norm,
                                    clip_norm_type=args.clip_norm_type,
                                    seed=getattr(args, 'random_seed', 3) * (self.rank + 1),
                                    world_size=self.world_size, first_env=first)

        self.session = Session()

        scope = getattr(self.network, 'name', 'local_learning')
        # tf.train.Saver() of actor_learner.py:79 is created AFTER apply_gradients, so the reference's checkpoints/ bundle holds
        # the RMSProp slots next to the network variables (the shipped -80000000.index files confirm it): written here too, so
        # that the reference's train.py can resume from a folder written by this implementation; optional on restore.
        self.network_saver = Saver(self._get_network_state, self._set_network_state, scope=scope,
                                   optional=lambda key: 'OptimizerVariables' in key)
        self.optimizer_saver = Saver(self._get_optimizer_state, self._set_optimizer_state, max_to_keep=1,
                                     name='OptimizerSaver', scope=scope)

    # ---- checkpoint payloads: TF variable names, reference layouts -----------------
    def _get_network_state(self):
        out = {n: t.detach().cpu().clone() for n, t in self.network.variables().items()}
        out.update(self._get_optimizer_state())
        return out

    def _set_network_state(self, state):
        for n, t in self.network.variables().items():
            t.copy_(state[n])
        self.network.params_changed()
        if all(k in state for k in self._get_optimizer_state()):
            self._set_optimizer_state(state)

    def _slot_views(self, flat):
        return {n: flat[off:off + int(np.prod(shape))].view(*shape) for n, off, shape, _ in self.network.tensors}

    def _get_optimizer_state(self):
        out = {}
        for n, t in self._slot_views(self.engine.ms).items():
            out[n + '/OptimizerVariables'] = t.detach().cpu().clone()
        for n, t in self._slot_views(self.engine.mom).items():
            out[n + '/OptimizerVariables_1'] = t.detach().cpu().clone()
        return out

    def _set_optimizer_state(self, state):
        for n, t in self._slot_views(self.engine.ms).items():
            t.copy_(state[n + '/OptimizerVariables'])
        for n, t in self._slot_views(self.engine.mom).items():
            t.copy_(state[n + '/OptimizerVariables_1'])

    def save_vars(self, force=False):
        if force or self.global_step - self.last_saving_step >= CHECKPOINT_INTERVAL:
            self.last_saving_step = self.global_step
            if self.rank == 0:
                self.network_saver.save(self.session, self.network_checkpoint_folder, global_step=self.last_saving_step)
                self.optimizer_saver.save(self.session, self.optimizer_checkpoint_folder, global_step=self.last_saving_step)

    def rescale_reward(self, reward):
        """ Clip immediate reward """
        if reward > 1.0:
            reward = 1.0
        elif reward < -1.0:
            reward = -1.0
        return reward

    def init_network(self):
        if not os.path.exists(self.network_checkpoint_folder):
            os.makedirs(self.network_checkpoint_folder, exist_ok=True)
        if not os.path.exists(self.optimizer_checkpoint_folder):
            os.makedirs(self.optimizer_checkpoint_folder, exist_ok=True)

        last_saving_step = self.network.init(self.network_checkpoint_folder, self.network_saver, self.session)

        path = Saver.latest_checkpoint(self.optimizer_checkpoint_folder)
        if path is not None:
            logging.info('Restoring optimizer variables from previous run')
            self.optimizer_saver.restore(self.session, path)

        if self.world_size > 1:      # replicas start identical; afterwards updates are bit-identical on all ranks
            torch.distributed.broadcast(self.network.params, src=0)
            self.network.params_changed()
            torch.distributed.broadcast(self.engine.ms, src=0)
            torch.distributed.broadcast(self.engine.mom, src=0)
        return last_saving_step

    def get_lr(self):
        if self.global_step <= self.lr_annealing_steps:
            return self.initial_lr - (self.global_step * self.initial_lr / self.lr_annealing_steps)
        else:
            return 0.0

    def cleanup(self):
        self.save_vars(True)
        self.session.close()
