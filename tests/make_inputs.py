"""Named seeded inputs shared by the golden-vector script and the tests."""
import numpy as np

import datagen

CASES = {
    "text_64k_l9": lambda: (datagen.text(65536, 101), 9, 65536),
    "text_300k_l9_nohint": lambda: (datagen.text(300_000, 102), 9, -1),
    "text_1p1M_l9_balanced": lambda: (datagen.text(1_100_000, 103), 9, 1_100_000),
    "mixed_1M_l9": lambda: (datagen.mixed(1_000_000, 125_000, 104), 9, 1_000_000),
    "random_200k_l4": lambda: (datagen.random_bytes(200_000, 105), 4, 200_000),
    "sparse_500k_l1": lambda: (datagen.sparse_binary(500_000, 106), 1, 500_000),
    "zeros_1M_l9": lambda: (np.zeros(1_000_000, np.uint8), 9, 1_000_000),
    "period3_90k_l9": lambda: (np.tile(np.frombuffer(b"abc", np.uint8), 30_000), 9, 90_000),
    "empty_l9": lambda: (np.zeros(0, np.uint8), 9, 0),
}
