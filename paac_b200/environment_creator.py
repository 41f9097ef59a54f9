"""EnvironmentCreator: mirror of environment_creator.py:1-15.

``args.game`` names an Atari ROM (needs ALE on the host) or ``synthetic`` (seeded raw 210x160 frames,
no emulator dependency).  The object exposes the same two things the learner uses: ``num_actions`` and
``create_environment(i)``.
"""


class EnvironmentCreator(object):

    def __init__(self, args):
        """
        Creates an object from which new environments can be created
        :param args:
        """
        if str(args.game).lower().startswith('synthetic'):
            from .synthetic_emulator import SyntheticEmulator
            self.num_actions = int(getattr(args, 'synthetic_actions', 6))
            self.create_environment = lambda i: SyntheticEmulator(i, args)
            return
        try:
            from .atari_emulator import AtariEmulator
            from ale_python_interface import ALEInterface
        except ImportError as e:
            raise ImportError("game %r needs the Arcade Learning Environment (ale_python_interface) on the host, "
                              "which is not installed here; use -g synthetic for seeded raw frames (%s)"
                              % (args.game, e))
        filename = args.rom_path + "/" + args.game + ".bin"
        ale_int = ALEInterface()
        ale_int.loadROM(str.encode(filename))
        self.num_actions = len(ale_int.getMinimalActionSet())
        self.create_environment = lambda i: AtariEmulator(i, args)
