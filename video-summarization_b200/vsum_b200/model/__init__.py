from .simnet import SimNet
from .simnet_pretrain import PretrainModel
