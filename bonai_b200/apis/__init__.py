from .train import (Trainer, set_random_seed, step_lr, init_dist, build_optimizer_args,
                    launcher_env)
