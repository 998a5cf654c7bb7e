"""B200-native drop-in for franka_test/scripts/control_torch (KL-ergodic controller).

Import paths match the reference (``control_torch.klerg``, ``.klerg_utils``,
``.barrier``, ``.dynamics``, ``.memory_buffer``, ``.default_policies``); all
arithmetic runs in libklerg_b200.so on the GPU.
"""
