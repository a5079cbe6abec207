"""The reference's constants (parameters.py:8-28); the values are part of the contract."""
C_PUCT = 5
EPS = 0.25
ALPHA = 0.2
PLAYOUT = 1600
DATA_DIR = "data"
MODEL_DIR = "models"
BATCH_SIZE = 2048
EPOCHS = 10
KL_TARG = 0.02
CHECK_FREQ = 10
LOG_LEVEL = 1
