"""PolicyVNetwork and the two composed classes: mirror of policy_v_network.py:6-64.

The actor softmax head, the critic head, the entropy term and the x5-scaled A2C loss are not graph nodes
here but kernels behind the C ABI (heads.cu, loss.cu); this class keeps the reference's attribute names as
fetch handles and creates the context once the trunk class has fixed the architecture.
"""
import numpy as np

from .networks import Network, NIPSNetwork, NatureNetwork, Placeholder, Fetch


class PolicyVNetwork(Network):

    def __init__(self, conf):
        """ Set up remaining layers, objective and loss functions (policy_v_network.py:8-57). """
        super(PolicyVNetwork, self).__init__(conf)

        self.entropy_regularisation_strength = conf['entropy_regularisation_strength']

        self.critic_target_ph = Placeholder('target', np.float32, [None])
        self.adv_actor_ph = Placeholder('advantage', np.float32, [None])

        # Final actor layer (softmax over num_actions) and final critic layer (linear, reshaped to [-1])
        self.output_layer_pi = Fetch(self, 'pi')
        self.output_layer_v = Fetch(self, 'v')
        self.log_output_layer_pi = Fetch(self, 'log_pi')
        self.output_layer_entropy = Fetch(self, 'entropy')
        self.critic_loss = Fetch(self, 'critic_loss')
        self.actor_objective_mean = Fetch(self, 'mean_actor_objective')
        self.critic_loss_mean = Fetch(self, 'mean_critic_loss')
        self.loss = Fetch(self, 'loss')

        self._create_context()


class NIPSPolicyVNetwork(PolicyVNetwork, NIPSNetwork):
    pass


class NaturePolicyVNetwork(PolicyVNetwork, NatureNetwork):
    pass
