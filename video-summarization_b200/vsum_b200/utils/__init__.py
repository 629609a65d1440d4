from .utils import set_seed, AverageMeter, load_yaml, load_json, mse_with_mask_loss
