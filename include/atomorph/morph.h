/*
 * atomorph/morph.h -- drop-in for the reference's morph.h (morph.h:11-76): class am::morph with
 * the reference's public members, signatures and error conventions (no exceptions; false /
 * nullptr / SIZE_MAX / transparent pixel on bad input).  Everything behind it -- blob detection,
 * blob matching, chain build, atom matching, rendering, fluid -- runs on one B200 through the
 * C-ABI of include/amx.h.  There is no CPU fallback: without a CUDA device synchronize() returns
 * false and get_pixels() yields transparent frames, and am::morph::last_error() says why.
 *
 * Threading contract as in the reference: one user thread; compute() /
 * iterate() are non-blocking and ignored while busy; setters are stored and reach the device at
 * the next successful synchronize(), which requires the worker to be paused (suspend()).
 */
#ifndef ATOMORPH_B200_MORPH_H
#define ATOMORPH_B200_MORPH_H

#include "atomorph.h"

namespace am {

class morph {
    public:
    morph();
    ~morph();

    void        clear              ();
    bool        add_pixel          (size_t frame, pixel px);
    bool        add_frame          (size_t frame);
    size_t      get_pixel_count    (size_t frame);
    size_t      get_frame_key      (double t);
    size_t      get_blob_count     (size_t frame);
    size_t      get_blob_count     ();
    pixel       get_average_pixel  (size_t frame);
    pixel       get_average_pixel  (size_t frame, size_t blob);
    pixel       get_pixel          (size_t frame, size_t position);
    color       get_background     (uint16_t x, uint16_t y, double t);
    point       interpolate        (point pt1, point pt2, double pt1_weight);
    pixel       interpolate        (pixel px1, pixel px2, double px1_weight);
    color       interpolate        (color c1, color c2, double c1_weight);
    color       interpolate        (color c1, color c2, double lag, double slope, double c1_weight);
    const blob* get_pixels         (size_t blob, double t, std::vector<pixel> *to);
    void        get_pixels         (double t, std::vector<pixel> *to);
    void        set_seed           (unsigned seed);
    double      get_time           (size_t current_frame, size_t total_frames);
    double      normalize_time     (double t);

    uint16_t    get_width       ();
    uint16_t    get_height      ();
    size_t      get_frame_count ();
    unsigned    get_state       ();
    void        next_state      ();
    const blob* get_blob        (size_t f, size_t b);
    double      get_energy      ();
    pixel       blob2pixel      (const blob *bl);
    void        compute         ();
    void        iterate         (size_t iterations);
    void        compute         (double seconds);
    void        suspend         ();
    bool        suspend         (double timeout);
    bool        is_busy         () const;

    bool        synchronize     ();

    void        set_blob_delimiter  (unsigned char d);
    void        set_blob_threshold  (double        t);
    void        set_blob_max_size   (size_t        s);
    void        set_blob_min_size   (size_t        s);
    void        set_blob_box_grip   (uint16_t      g);
    void        set_blob_box_samples(size_t        s);
    void        set_blob_number     (size_t        n);
    void        set_blob_rgba_weight(unsigned char w);
    void        set_blob_size_weight(unsigned char w);
    void        set_blob_xy_weight  (unsigned char w);
    void        set_degeneration    (size_t        d);
    void        set_motion          (unsigned char m);
    void        set_fading          (unsigned char f);
    void        set_threads         (size_t        t);
    void        set_cycle_length    (size_t        c);
    void        set_feather         (size_t        f);
    void        set_keep_background (bool          k);
    void        set_finite          (bool          f);
    void        set_show_blobs      (unsigned      b);

    // These restart the morph (identifier change), as in the reference:
    void        set_fluid           (unsigned      f);
    void        set_density         (uint16_t      d);

    void        set_resolution      (uint16_t w, uint16_t h);

    // ---- additions (not in the reference) ----
    const char* last_error      () const;           // why the last device call failed ("" if none)
    void*       device_context  ();                 // the amx_ctx* underneath (include/amx.h), for batch rendering

    private:
    morph(const morph &);
    morph &operator=(const morph &);
    struct impl;
    impl *p;
};

}

#endif
