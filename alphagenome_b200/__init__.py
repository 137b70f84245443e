"""alphagenome_b200 — B200-native (sm_100a) AlphaGenome trunk forward behind the reference's module API.

    from alphagenome_b200 import AlphaGenome, install
"""
from .model import AlphaGenome, Embeds, TransformerUnetOutput, publication_heads_config, set_update_running_var  # noqa: F401
from .install import install  # noqa: F401
