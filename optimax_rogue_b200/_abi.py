"""ctypes mirror of ``include/orx.h`` (struct layouts, codes, function prototypes).

Kept free of torch so that the CPU test-suite can check the ABI without a GPU.
"""
import ctypes as C

ABI_VERSION = 5

MAX_NPC = 8
MAX_EVENTS_BASE = 4
MAX_DIM = 255
NO_STAIRS = 255
OBS_LEN = 12
STAT_COUNT = 8
STAT_NAMES = ('ticks', 'p1_wins', 'p2_wins', 'ties', 'events', 'descents', 'hits', 'reserved')

DGEN_EMPTY, DGEN_FIXED = 0, 1
START_TOGETHER, START_SEPARATED = 0, 1
BOT_NONE, BOT_RANDOM, BOT_STAIRCASE = 0, 1, 2
EV_NONE, EV_MOVE, EV_COMBAT, EV_DUNGEON, EV_DEATH, EV_DESCEND = 0, 1, 2, 3, 4, 5

OK, ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_CUDA_BASE = 0, -1, -2, -100

# OrxConfig.path_flags (include/orx.h)
PATH_NO_TENSOR_MAP, PATH_NO_NPC_PIPE, PATH_STATIC_TILES, PATH_NO_EVENT_PIPE, PATH_HOST_STAGED, PATH_TILE_FLAGS = 1, 2, 4, 8, 16, 32
PATH_TILES_PER_CTA_SHIFT = 8
SCHED_HEADER_WORDS = 4
TILE = 256


def cmd5_bytes(n: int) -> int:
    return (5 * int(n) + 7) // 8


def res2_bytes(n: int) -> int:
    return (2 * int(n) + 7) // 8


def sched_words(n: int) -> int:
    """orx_sched_words: scratch words that enable tile-by-tile ordering for n games."""
    return SCHED_HEADER_WORDS + 2 * (max(int(n), 0) // TILE)


class OrxConfig(C.Structure):
    _fields_ = [
        ('struct_size', C.c_uint32),
        ('width', C.c_int32), ('height', C.c_int32),
        ('dgen_kind', C.c_int32),
        ('start_kind', C.c_int32),
        ('start_depth', C.c_int32 * 2),
        ('despawn_strat', C.c_int32),
        ('max_ticks', C.c_int32),
        ('hp', C.c_int32 * 2),
        ('damage', C.c_int32 * 2),
        ('armor', C.c_int32 * 2),
        ('auto_reset', C.c_int32),
        ('n_npc', C.c_int32),
        ('seed', C.c_uint64),
        ('fixed_tiles', C.c_void_p),
        ('fixed_ground', C.c_void_p),
        ('fixed_n_ground', C.c_int32),
        ('fixed_stairs', C.c_int32 * 2),
        ('path_flags', C.c_uint32),
    ]


class OrxState(C.Structure):
    _fields_ = [
        ('pos', C.c_void_p), ('hp', C.c_void_p), ('depth', C.c_void_p), ('stairs', C.c_void_p),
        ('tick', C.c_void_p), ('episode', C.c_void_p), ('status', C.c_void_p),
        ('npc_pos', C.c_void_p), ('npc_hp', C.c_void_p), ('npc_depth', C.c_void_p),
        ('sched', C.c_void_p), ('sched_words', C.c_uint32), ('reserved', C.c_uint32),
        ('flat', C.c_void_p),
    ]


class OrxEvent(C.Structure):
    _fields_ = [('kind', C.c_uint8), ('iden', C.c_uint8), ('a', C.c_uint8), ('b', C.c_uint8),
                ('depth', C.c_int32)]


# ---- ruleset R1 (README-only rules; docs/RULESET_R1.md) ------------------------------------------
R1_LANES, R1_ENEMIES, R1_ITEMS, MOVE_HEAL, R1_STATE_BYTES, R1_OBS_LEN = 16, 8, 4, 6, 241, 64
EV_SPAWN, EV_HEALTH, EV_PICKUP, EV_XP = 6, 7, 8, 9                       # R1 replication-log kinds (include/orx.h)
R1_HIT_FULL, R1_HIT_HALF, R1_HIT_NEGATED, R1_HIT_CONTEST = 1, 2, 3, 4
R1_HEALTH_HEAL, R1_HEALTH_SEPARATION = 1, 2
R1_MAX_EVENTS = 64


class OrxR1Config(C.Structure):
    _fields_ = [('struct_size', C.c_uint32), ('width', C.c_int32), ('height', C.c_int32),
                ('max_ticks', C.c_int32), ('auto_reset', C.c_int32), ('wall_density', C.c_int32),
                ('seed', C.c_uint64), ('path_flags', C.c_uint32), ('reserved', C.c_uint32)]


R1_PATH_HALFWARP, R1_PATH_BLOCK_FLAGS, R1_BLOCK = 1, 2, 128


def r1_sched_words(n: int) -> int:
    """ORX_R1_SCHED_WORDS(n), sized for 64-game blocks so that tuning builds with smaller CTAs fit as well."""
    return 2 * ((int(n) + 63) // 64)


R1_PLANES = (('ent_loc', 'int32', (16,)), ('ent_depth', 'int32', (16,)), ('ent_stat', 'int32', (16,)),
             ('pl_a', 'int32', (2,)), ('pl_b', 'int32', (2,)), ('pl_c', 'int32', (2,)),
             ('lvl_stairs', 'int32', ()), ('lvl_key', 'int32', (2,)), ('sep', 'int32', ()),
             ('tick', 'int32', ()), ('episode', 'int32', ()), ('status', 'uint8', ()))


class OrxR1State(C.Structure):
    _fields_ = [(name, C.c_void_p) for name, _, _ in R1_PLANES] + [('sched', C.c_void_p), ('sched_words', C.c_uint32),
                                                                    ('reserved', C.c_uint32)]


# name -> (restype, argtypes); every symbol include/orx.h declares
PROTOTYPES = {
    'orx_abi_version': (C.c_int, []),
    'orx_strerror': (C.c_char_p, [C.c_int]),
    'orx_sched_words': (C.c_size_t, [C.c_int64]),
    'orx_state_bytes': (C.c_size_t, [C.POINTER(OrxConfig)]),
    'orx_max_events': (C.c_int, [C.POINTER(OrxConfig)]),
    'orx_event_count_add': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    'orx_reset': (C.c_int, [C.POINTER(OrxConfig), C.POINTER(OrxState), C.c_void_p, C.c_int,
                            C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_step': (C.c_int, [C.POINTER(OrxConfig), C.POINTER(OrxState), C.c_void_p, C.c_void_p,
                           C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_step_packed': (C.c_int, [C.POINTER(OrxConfig), C.POINTER(OrxState), C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_step_host_packed': (C.c_int, [C.POINTER(OrxConfig), C.POINTER(OrxState), C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_step_host_packed_sync': (C.c_int, [C.POINTER(OrxConfig), C.POINTER(OrxState), C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_step_host': (C.c_int, [C.POINTER(OrxConfig), C.POINTER(OrxState), C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_step_host_sync': (C.c_int, [C.POINTER(OrxConfig), C.POINTER(OrxState), C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_step_bits': (C.c_int, [C.POINTER(OrxConfig), C.POINTER(OrxState), C.c_void_p, C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_step_host_bits': (C.c_int, [C.POINTER(OrxConfig), C.POINTER(OrxState), C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_step_host_bits_sync': (C.c_int, [C.POINTER(OrxConfig), C.POINTER(OrxState), C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_bot_moves': (C.c_int, [C.POINTER(OrxConfig), C.POINTER(OrxState), C.c_int, C.c_int,
                                C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_rollout': (C.c_int, [C.POINTER(OrxConfig), C.POINTER(OrxState), C.c_int, C.c_int, C.c_int,
                              C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_replay': (C.c_int, [C.POINTER(OrxConfig), C.POINTER(OrxState), C.c_void_p, C.c_void_p, C.c_int,
                             C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_step_bots': (C.c_int, [C.POINTER(OrxConfig), C.POINTER(OrxState), C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_step_observe': (C.c_int, [C.POINTER(OrxConfig), C.POINTER(OrxState), C.c_void_p, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_observe': (C.c_int, [C.POINTER(OrxConfig), C.POINTER(OrxState), C.c_void_p, C.c_int,
                              C.c_int64, C.c_void_p]),
    'orx_observe_npc': (C.c_int, [C.POINTER(OrxConfig), C.POINTER(OrxState), C.c_void_p, C.c_int64, C.c_void_p]),
    'orx_r1_reset': (C.c_int, [C.POINTER(OrxR1Config), C.POINTER(OrxR1State), C.c_void_p, C.c_int,
                               C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_r1_step': (C.c_int, [C.POINTER(OrxR1Config), C.POINTER(OrxR1State), C.c_void_p, C.c_void_p,
                              C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_r1_rollout': (C.c_int, [C.POINTER(OrxR1Config), C.POINTER(OrxR1State), C.c_int, C.c_void_p,
                                 C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_r1_observe': (C.c_int, [C.POINTER(OrxR1Config), C.POINTER(OrxR1State), C.c_void_p, C.c_int,
                                 C.c_int64, C.c_void_p]),
    'orx_r1_step_events': (C.c_int, [C.POINTER(OrxR1Config), C.POINTER(OrxR1State), C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_r1_bot_moves': (C.c_int, [C.POINTER(OrxR1Config), C.POINTER(OrxR1State), C.c_int, C.c_int, C.c_void_p,
                                   C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_r1_replay': (C.c_int, [C.POINTER(OrxR1Config), C.POINTER(OrxR1State), C.c_void_p, C.c_void_p, C.c_int,
                                C.c_int64, C.c_uint64, C.c_void_p]),
    'orx_r1_step_host_sync': (C.c_int, [C.POINTER(OrxR1Config), C.POINTER(OrxR1State), C.c_void_p, C.c_void_p,
                                        C.c_int64, C.c_uint64, C.c_void_p]),
}


def bind(lib, prototypes=None):
    """Sets restype/argtypes on a loaded CDLL; raises AttributeError on a missing symbol."""
    for name, (res, args) in (prototypes or PROTOTYPES).items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
