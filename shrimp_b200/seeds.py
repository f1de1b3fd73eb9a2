"""Spaced seeds: host mirror of gmapper/seeds.c (add_spaced_seed :9-42, load_default_seeds :56-80,
load_default_mirna_seeds :45-50) and the default tables of gmapper-defaults.h:194-238."""
from __future__ import annotations

from dataclasses import dataclass

MAX_SEED_SPAN = 64          # gmapper-definitions.h:55
MAX_SEED_WEIGHT = 14        # without -H
HASH_TABLE_POWER = 12

# gmapper-defaults.h:197-238: index = weight - 10 (letter space and colour space tables are equal)
_DEFAULT_SEEDS = {
    10: ["111110011111", "111100110001111", "111100100100100111", "111001000100001001111"],
    11: ["1111001111111", "1111100110001111", "11110010010001001111", "11100110010000100100111"],
    12: ["11110111101111", "1111011100100001111", "1111000011001101111"],
    16: ["111111101110111111", "1111100101101101011111", "11110011001010100011011111",
         "111101001100000100110011010111"],
    18: ["11111011111110111111", "11110111011010111011111", "11111100110101101001011111",
         "11111010101100100010011101111"],
}
_MIRNA_SEEDS = ["00111111001111111100", "00111111110011111100", "00111111111100111100",
                "00111111111111001100", "00111111111111110000"]
DEFAULT_WEIGHT = 12


@dataclass(frozen=True)
class Seed:
    mask: int     # bit 0 = rightmost character of the seed string
    span: int
    weight: int
    string: str


def add_spaced_seed(seed_string: str) -> Seed:
    """seeds.c:9-42.  Raises ValueError where the reference returns false."""
    span = len(seed_string)
    weight = seed_string.count("1")
    if span < 1 or span > MAX_SEED_SPAN or weight < 1 or seed_string.count("0") != span - weight:
        raise ValueError(f"invalid spaced seed [{seed_string}]")
    mask = 0
    for ch in seed_string:          # bitmap_prepend: earlier characters end up in higher bits
        mask = (mask << 1) | (1 if ch == "1" else 0)
    return Seed(mask, span, weight, seed_string)


def load_default_seeds(weight: int = 0) -> list[Seed]:
    """seeds.c:56-80 (same tables for letter and colour space)."""
    if weight == 0:
        weight = DEFAULT_WEIGHT
    if weight < 10 or weight > 18 or weight not in _DEFAULT_SEEDS:
        raise ValueError(f"no default seeds of weight {weight}")
    return [add_spaced_seed(s) for s in _DEFAULT_SEEDS[weight]]


def load_default_mirna_seeds() -> list[Seed]:
    return [add_spaced_seed(s) for s in _MIRNA_SEEDS]


def parse_seeds(spec: str) -> list[Seed]:
    """-s option: comma separated seed strings, or w<weight> for a default set (gmapper.c:1830-1850)."""
    out: list[Seed] = []
    for tok in spec.split(","):
        tok = tok.strip()
        if tok.startswith("w"):
            out += load_default_seeds(int(tok[1:]))
        else:
            out.append(add_spaced_seed(tok))
    return out
