"""heal_swin_b200 -- B200-native (sm_100a) engine for the HEAL-SWIN forward/backward hot path.

Drop-in module surface: ``heal_swin_b200.models_torch`` mirrors
``heal_swin.models_torch`` (hp_windowing, hp_shifting, swin_hp_transformer).
All device work goes through the C-ABI library ``libhealswin_b200.so``
(``include/healswin_b200.h``); there is no CPU fallback for device ops.
"""
__version__ = "0.1.0"
