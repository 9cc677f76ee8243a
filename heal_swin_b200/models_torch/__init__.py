"""Drop-in mirror of heal_swin.models_torch for the HEAL-SWIN hot path."""
