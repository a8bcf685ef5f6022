"""csm_hf_b200: B200-native CSM frame generation behind the reference's CSMModel API."""
from .config import CSMConfig, CSMOutput, LlamaDims, tiny_config  # noqa: F401
