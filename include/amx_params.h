/*
 * amx_params.h -- parameter ids shared by the C-ABI (amx_morph_set) and the
 * oracle harness (amref_set).  One id per am::morph setter of the reference
 * (reference morph.h:52-76).  Values travel as double; ids are stable ABI.
 */
#ifndef AMX_PARAMS_H
#define AMX_PARAMS_H

enum amx_param {
    AMX_P_BLOB_DELIMITER   = 0,  /* morph.h:52  set_blob_delimiter  (am::RGB | am::HSP)      */
    AMX_P_BLOB_THRESHOLD   = 1,  /* morph.h:53  set_blob_threshold  (0..1)                    */
    AMX_P_BLOB_MAX_SIZE    = 2,  /* morph.h:54  set_blob_max_size   (>=1.8e19 => SIZE_MAX)    */
    AMX_P_BLOB_MIN_SIZE    = 3,  /* morph.h:55  set_blob_min_size                             */
    AMX_P_BLOB_BOX_GRIP    = 4,  /* morph.h:56  set_blob_box_grip                             */
    AMX_P_BLOB_BOX_SAMPLES = 5,  /* morph.h:57  set_blob_box_samples                          */
    AMX_P_BLOB_NUMBER      = 6,  /* morph.h:58  set_blob_number                               */
    AMX_P_BLOB_RGBA_WEIGHT = 7,  /* morph.h:59  set_blob_rgba_weight                          */
    AMX_P_BLOB_SIZE_WEIGHT = 8,  /* morph.h:60  set_blob_size_weight                          */
    AMX_P_BLOB_XY_WEIGHT   = 9,  /* morph.h:61  set_blob_xy_weight                            */
    AMX_P_DEGENERATION     = 10, /* morph.h:62  set_degeneration                              */
    AMX_P_DENSITY          = 11, /* morph.h:75  set_density (restarts the morph)              */
    AMX_P_MOTION           = 12, /* morph.h:63  set_motion  (am::NONE|LINEAR|SPLINE)          */
    AMX_P_FADING           = 13, /* morph.h:64  set_fading  (am::NONE|LINEAR|COSINE|PERLIN)   */
    AMX_P_THREADS          = 14, /* morph.h:65  set_threads                                   */
    AMX_P_CYCLE_LENGTH     = 15, /* morph.h:66  set_cycle_length                              */
    AMX_P_FEATHER          = 16, /* morph.h:67  set_feather                                   */
    AMX_P_KEEP_BACKGROUND  = 17, /* morph.h:68  set_keep_background                           */
    AMX_P_FINITE           = 18, /* morph.h:69  set_finite                                    */
    AMX_P_SHOW_BLOBS       = 19, /* morph.h:70  set_show_blobs (am::TEXTURE|AVERAGE|DISTINCT) */
    AMX_P_FLUID            = 20, /* morph.h:74  set_fluid   (restarts the morph)              */
    AMX_P_SEED             = 21, /* morph.h:31  set_seed                                      */
    AMX_P_COUNT_
};

#endif
