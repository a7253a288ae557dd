"""Constants of the caption-inference hot path.

Mirror of the reference's star-imported config module
(/root/reference/common/common_definitions.py:6-70): same names, same values, no
TensorFlow.  `ACTIVATION` / `KERNEL_INITIALIZER` are tf callables in the reference
(common_definitions.py:14-15); here they are the strings the engine understands.
"""

IS_TRAINING = False          # inference-only build (reference default True, common_definitions.py:6)
USE_GPU = True               # common_definitions.py:8 ; there is no CPU fallback in this build

TOP_K = 10000                # tokenizer vocabulary cap, common_definitions.py:12

ACTIVATION = "leaky_relu"    # tf.nn.leaky_relu, alpha=0.2 (TF default), common_definitions.py:14
LEAKY_RELU_ALPHA = 0.2
KERNEL_INITIALIZER = "he_normal"  # common_definitions.py:15

IMAGE_INPUT_SIZE = 512       # common_definitions.py:18
BATCH_SIZE = 10
BEAM_SEARCH_N = 4            # common_definitions.py:22 (README quotes BEAM_SIZE=8 for the published scores)
N_VAL_DATASET = 50
AMOUNT_OF_VALIDATION = 100
DROPOUT_RATE = 0.1           # identity at inference

TOKENIZER_FILENAME = "datasets/_tokenizer.json"
ADDITIONAL_FILENAME = "datasets/_additional_extractor.json"
RETINANET_WEIGHT_PATH = "model_weights/mobilenet224_1.0_coco.h5"
TRANSFORMER_WEIGHT_PATH = "model_weights/multimodal_transformer.h5"
TRANSFORMER_CHECKPOINT_PATH = "./checkpoints/train/multimodal_transformer"

# Transformer hyper-parameters, common_definitions.py:56-59
num_layers = 6
d_model = 512
dff = 2048
num_heads = 8

# RetinaNet / FPN parameters, common_definitions.py:63-67
NUM_OF_CLASSES = 80
NUM_OF_RETINANET_FILTERS = 256
NUM_OF_ANCHORS = 9
NUM_OF_PYRAMIDS = 5
N_CONV_SUBMODULE = 2

# MT encoder, common_definitions.py:70
BASELINE_INDEX = 3

LAYERNORM_EPS = 1e-6         # transformer.py:170-171,216-218,264

# Synthetic-benchmark conventions (SURVEY.md §8d / BASELINE.md §2); not in the reference.
SYNTH_VOCAB = TOP_K
SYNTH_MAX_SEQ_LEN = 64
PAD_ID, UNK_ID, START_ID, END_ID = 0, 1, 2, 3

BACKBONES = ("mobilenet224_1.0", "resnet50", "densenet121")
