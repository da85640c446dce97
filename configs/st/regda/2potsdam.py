"""Hyper-parameters of st.regda.2potsdam (reference configs/st/regda/2potsdam.py:6-48, configs/ToPotsdam.py:58): the
attribute names the trainer reads are the reference's.  The reference's *_DATA_CONFIG dicts describe file lists and
albumentations pipelines (CPU data loading, out of scope here); SYNTHETIC describes the seeded synthetic tensors of the
same shapes that tools/train_ssl_reg.py --data synthetic trains on."""
MODEL = 'ResNet101'
IGNORE_LABEL = -1
CLASS_NUM = 6                # len(IsprsDA.LABEL_MAP)
MOMENTUM = 0.9
SNAPSHOT_DIR = './log/regda/2potsdam'
TARGET_SET = 'Potsdam'
DATASETS = 'IsprsDA'        # class 0 is dropped from the mIoU (regda/utils/eval.py:16-17)
WEIGHT_DECAY = 0.0005
LEARNING_RATE = 1e-2
STAGE1_STEPS = 4000
STAGE2_STEPS = 6000
STAGE3_STEPS = 6000
NUM_STEPS = None             # for learning rate poly (set by the trainer: 1.5 * STAGE3_STEPS)
PREHEAT_STEPS = None         # for warm-up            (set by the trainer: STAGE3_STEPS / 20)
POWER = 0.9
EVAL_EVERY = 500
GENE_EVERY = 1000
CUTOFF_TOP = 0.8
CUTOFF_LOW = 0.6
BATCH_SIZE = 8               # per domain, per GPU
SYNTHETIC = dict(size=(512, 512), regions_per_tile=200)
