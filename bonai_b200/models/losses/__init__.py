from .losses import (CrossEntropyLoss, L1Loss, SmoothL1Loss, FocalLoss, accuracy, Accuracy,
                     weight_reduce_loss, reduce_loss)

__all__ = ['CrossEntropyLoss', 'L1Loss', 'SmoothL1Loss', 'FocalLoss', 'accuracy', 'Accuracy',
           'weight_reduce_loss', 'reduce_loss']
