"""Training and model configuration: the constants of reference src/config.py:7-42 (class attributes that the
train scripts mutate at run time, train_rna2dna.py:167-174).  Not accelerated; part of the boundary."""
import torch


class Config:
    # model geometry (overridable through INPUT_DIM_A / INPUT_DIM_B / LATENT_DIM in the train scripts)
    INPUT_DIM_A = 1177
    INPUT_DIM_B = 1211
    LATENT_DIM = 20

    # optimisation
    BATCH_SIZE = 32
    NUM_EPOCHS = 200
    LEARNING_RATE = 5e-4
    WEIGHT_DECAY = 1e-5

    # loss weights: beta = min(1, epoch / BETA_WARMUP_EPOCHS) * BETA_START
    BETA_START = 1e-3
    BETA_WARMUP_EPOCHS = 50
    GAMMA = 1.0

    PATIENCE = 15
    LR_SCHEDULER_FACTOR = 0.5
    LR_SCHEDULER_PATIENCE = 5

    CHECKPOINT_DIR = 'checkpoints'
    BEST_MODEL_NAME = 'best_multivae.pt'

    # This implementation exists for CUDA (sm_100a); the other choices are kept so that importing the
    # config on a machine without a GPU behaves like the reference.
    DEVICE = torch.device("cuda" if torch.cuda.is_available()
                          else "mps" if torch.backends.mps.is_available() else "cpu")

    TRAIN_TEST_SPLIT = 0.2
    RANDOM_SEED = 42
