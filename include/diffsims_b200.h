/*
 * diffsims_b200 -- C ABI of the B200-native kinematical template-simulation path.
 *
 * The reference (pyxem/diffsims 0.7.0) is pure Python and has no FFI seam; the
 * boundary it offers is its Python API.  The Python
 * mirror of that API in diffsims_b200/ reaches the sm_100a kernels ONLY through
 * the entry points below (ctypes, tensor.data_ptr()).  INTEGRATION.md shows the
 * ctypes stub a reference maintainer would add at each cited call site.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host;
 *  - the library never allocates or frees device memory and never synchronises:
 *    work is enqueued on `stream` (a cudaStream_t, may be NULL = legacy stream);
 *  - return value 0 = ok; negative = error (ds_last_error() gives the text);
 *  - no C++ exceptions cross the boundary; no torch types appear here;
 *  - doubles are IEEE binary64, row-major arrays, sizes in elements.
 *
 * All reference citations are relative to /root/reference.
 */
#ifndef DIFFSIMS_B200_H
#define DIFFSIMS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DS_ABI_VERSION 2

/* atomic scattering factor parameterisations, diffsims/utils/sim_utils.py:227-253 */
#define DS_SCATT_NONE    0 /* f = 1 for every atom (sim_utils.py:284-286)          */
#define DS_SCATT_LOBATO  1 /* sum a_i (2 + b_i g^2) / (1 + b_i g^2)^2              */
#define DS_SCATT_XTABLES 2 /* sum a_i exp(-b_i g^2 / 4)                             */

/* rel-rod shape factors, diffsims/utils/shape_factor_models.py */
#define DS_SHAPE_BINARY      0 /* :33-49   */
#define DS_SHAPE_LINEAR      1 /* :52-73   */
#define DS_SHAPE_SINC        2 /* :76-101  (0 at s == 0, as the reference's where= branch) */
#define DS_SHAPE_SIN2C       3 /* :104-123 */
#define DS_SHAPE_ATANC       4 /* :126-151 (1 at s == 0) */
#define DS_SHAPE_LORENTZIAN  5 /* :154-180 */
#define DS_SHAPE_LORENTZIAN_PRECESSION 6 /* :183-219, chosen when precession != 0 and approximate */
#define DS_SHAPE_NONE_RETURN_S 7 /* prefactor 1, no threshold: caller applies a Python callable */

int ds_abi_version(void);
const char *ds_last_error(void);

/*
 * Process-wide tuning options (schedules only, never results).  Each option also has an environment variable
 * that is read ONCE, at the first use of the library; launch paths afterwards only do atomic loads, so
 * launches on different streams / host threads never race.  value -1 = "not set" (the measured default).
 *   render_pipe (DS_RENDER_PIPE) 0: K3 without the warp-specialised pipelined kernel
 *   render_group (DS_RENDER_GROUP) 1/2/4/8: warps per template of the phase-synchronous K3 kernel
 *   render_fronts (DS_RENDER_FRONTS) 1..3: preparation warps of the pipelined K3 kernel
 *   render_pipe_maxcap (DS_RENDER_PIPE_MAXCAP): largest row capacity the pipelined K3 kernel takes
 *   render_nostage (DS_RENDER_NOSTAGE) 1: K3 reads spot rows in place instead of cp.async.bulk staging
 *   render_mma (DS_RENDER_MMA) 0/1, render_mma_min, render_mma_tmpl_min: legacy mma.sync path of the
 *       phase-synchronous / pipelined kernels and its thresholds
 *   render_umma (DS_RENDER_UMMA) 0/1: tcgen05 K3 kernel off / forced (default: by capacity and density hint)
 *   render_umma_window (DS_RENDER_UMMA_WINDOW) 0: tcgen05 K3 kernel without per-chunk column windows
 *   render_umma_team (DS_RENDER_UMMA_TEAM) 4: tcgen05 per-reflection K3 kernel with four producer warps per stage
 *   render_rows (DS_RENDER_ROWS) 0/1: row-binned tcgen05 K3 kernel off / forced (default: >= 320 reflections per template)
 *   render_rows_stages (DS_RENDER_ROWS_STAGES) 3..6: operand stages of the row-binned kernel
 *   render_zero_tma (DS_RENDER_ZERO_TMA) 1: pipelined K3 kernel stores all-zero regions through the TMA engine
 *   sim_lines (DS_SIM_LINES) 0/1: K2 scan-line cull off / forced
 *   sim_split (DS_SIM_SPLIT) 1/2/4/8: warps (= rotations) per CTA of the warp-per-rotation K2 kernel
 *   sim_cta (DS_SIM_CTA) 0/1: CTA-per-rotation K2 kernel off / forced (default: large tables, or few rotations over >= 4096 rows)
 *   sim_stash (DS_SIM_STASH) n: candidate capacity of a rotation in the CTA-per-rotation K2 kernel (default 4096)
 * Thread safety of the library: entry points may be called concurrently from several host threads as long
 * as the calls use different streams and different output buffers (and different `ticket` words for
 * ds_render); one process may drive several devices.
 */
int ds_set_option(const char *name, int32_t value);
int ds_get_option(const char *name, int32_t *value);

/*
 * K1 -- kinematical structure factors, once per (phase, g-set).
 * Replaces _get_kinematical_structure_factor (diffsims/utils/sim_utils.py:256-304),
 * get_atomic_scattering_factors (:227-253) and the |F|^2 of
 * get_kinematical_intensities (:307-354).
 *
 *   F(g) = sum_e f_e(g^2) exp(-g^2 B_e / 4) sum_{j in e} occ_j exp(2 pi i hkl . r_j)
 *
 * Atoms are grouped by element: atoms elem_start[e] .. elem_start[e+1]-1 belong to
 * element e.  frac holds the fractional coordinates ALREADY multiplied by
 * inv(stdbase @ recbase) (sim_utils.py:290-291; a 3x3 host-side matmul).
 * Outputs (either may be NULL): F_out[n_g][2] = (re, im); I_out[n_g] =
 * prefactor[g] * |F|^2 (prefactor NULL = 1).
 * hkl_int_max > 0 is the caller's promise that every entry of hkl is an integer with |.| <= hkl_int_max (<= 127): with
 * a scratch buffer of ds_structure_factors_scratch_bytes(n_atoms, hkl_int_max) bytes (16-byte aligned) large cells then
 * take the factorised kernels (per-atom phase tables Ex[h] Ey[k] Ez[l], two complex multiplications per atom x g pair
 * instead of a sincospi).  hkl_int_max = 0 / table_scratch = NULL: the direct evaluation (any real hkl).
 */
int64_t ds_structure_factors_scratch_bytes(int32_t n_atoms, int32_t hkl_int_max);
int ds_structure_factors(void *stream,
                         int32_t n_g, const double *hkl /*[n_g][3]*/, const double *gnorm /*[n_g]*/,
                         int32_t n_atoms, const double *frac /*[n_atoms][3]*/, const double *occ /*[n_atoms]*/,
                         int32_t n_elem, const int32_t *elem_start /*[n_elem+1]*/,
                         const double *coeffs /*[n_elem][5][2] (a_i, b_i)*/, const double *dw /*[n_elem]*/,
                         int32_t scattering_model,
                         const double *prefactor /*[n_g] or NULL*/,
                         double *F_out /*[n_g][2] or NULL*/, double *I_out /*[n_g] or NULL*/,
                         int32_t hkl_int_max, void *table_scratch /* or NULL */);

/*
 * Pack the per-phase g table for K2: out[g] = (gx, gy, gz, |g|^2) as float (16-byte rows,
 * the tile format K2 stages into shared memory with cp.async.bulk).
 *
 * Optional extinction marking (g_I0 = NULL, ref_row < 0 or rel_cut <= 0: none): rows with
 * g_I0[g] <= rel_cut * g_I0[ref_row] get |g|^2 = +inf, which the float32 cull of ds_simulate rejects, so the
 * systematically absent reflections of centred lattices (3/4 of an F lattice, 13/16 of diamond) cost no float64
 * work.  This is exact when (i) ref_row is the direct beam (000), which every rotation excites with s = 0,
 * (ii) the shape factor obeys sf(s) <= sf(0) (binary, linear, atanc, lorentzian without precession) and
 * (iii) rel_cut <= min_intensity / 2: such a row has I = sf(s) I0 <= rel_cut sf(0) I0[000] < max(I) min_intensity
 * and the reference's cut I > max(I) * minimum_intensity (simulation_generator.py:237-241) removes it anyway.
 */
int ds_pack_gtable(void *stream, int32_t n_g, const double *g_xyz /*[n_g][3]*/, float *g_f32 /*[n_g][4]*/,
                   const double *g_I0 /*[n_g] or NULL*/, int32_t ref_row, double rel_cut);

/*
 * K2 -- fused rotate / excitation error / shape factor / cull / threshold over
 * (rotation x g).  Replaces the body of the rotation loop of
 * SimulationGenerator.calculate_diffraction2d (diffsims/generators/simulation_generator.py:211-241):
 * DiffractingVector.rotate_with_basis (crystallography/_diffracting_vector.py:127-161),
 * get_intersecting_reflections (simulation_generator.py:319-412), the prefactor * |F|^2 of
 * get_kinematical_intensities (utils/sim_utils.py:353) and the minimum_intensity threshold (:237);
 * and likewise DiffractionGenerator.calculate_ed_data (generators/diffraction_generator.py:247-324).
 *
 *   g_lab = R(q) g,  R(q) the active rotation matrix of the unit quaternion q = (a, b, c, d)
 *   s     = (r_s - sqrt(r_s^2 - x^2 - y^2)) - z,   r_s = inv_wavelength
 *   keep  |s| < s_max (strict)           [precession: the two-surface test of :365-375]
 *   precession_rad != 0 with DS_SHAPE_LORENTZIAN_PRECESSION uses the closed form (:183-219); with any other
 *   model that model is averaged numerically over the precession circle (_shape_factor_precession, :222-269)
 *   I     = shape(s; width) * I0[g];     keep I > max_rot(I) * min_intensity   (min_intensity < 0: keep all)
 *
 * Output is padded per rotation: row r holds count[r] reflections in g-table order,
 * at most `cap`; *max_count (device int32, caller zero-initialised) receives the largest
 * number of reflections any rotation needed, so the caller can retry with cap >= *max_count.
 * excitation_error may be NULL.
 *
 * Optional scan-line description of the table (n_lines = 0 / NULL pointers = none): rows line_start[L] ..
 * line_start[L+1]-1 of the table are the lattice line line_g0[L] + i * line_step (i = 0, 1, ...), i.e. consecutive
 * Miller indices along one axis, as both g-set enumerations of the reference produce them.  With it the coarse cull
 * may solve for the short index interval in which each line crosses the Ewald slab instead of testing every row
 * (same candidates, same order, same float64 refine; used for tables of >= 2048 rows when the slab is thinner than
 * half a step, DS_SIM_LINES=0/1 in the environment forces it off/on).  line_g0: device float[n_lines][4] (16-byte aligned),
 * line_start: device int32[n_lines+1] padded to a multiple of 4 entries, line_step_host: HOST double[3].
 */
int ds_simulate(void *stream,
                int32_t n_rot, const double *quat /*[n_rot][4]*/,
                int32_t n_g, const double *g_xyz /*[n_g][3]*/, const float *g_f32 /*[n_g][4]*/,
                const double *g_I0 /*[n_g]*/,
                double g_max /* upper bound of |g| in the table (reciprocal radius) */,
                double inv_wavelength, double s_max, double width,
                int32_t shape_model, double minima_number, double precession_rad,
                double min_intensity,
                int32_t cap,
                int32_t *count /*[n_rot]*/, int32_t *g_index /*[n_rot][cap]*/,
                double *xyz /*[n_rot][cap][3]*/, double *intensity /*[n_rot][cap]*/,
                double *excitation_error /*[n_rot][cap] or NULL*/,
                int32_t *max_count /*[1]*/,
                int32_t n_lines, const float *line_g0, const int32_t *line_start, const double *line_step_host);

/*
 * K3 -- rasterise spot lists into templates.
 * Replaces Simulation2D._get_transformed_coordinates (diffsims/simulations/simulation2d.py:261-285),
 * the in-frame test / truncation / normalisation of get_diffraction_pattern (:357-442) and
 * get_pattern_from_pixel_coordinates_and_intensities (diffsims/pattern/detector_functions.py:251-311):
 *   fast == 1: integer pixels, last-write-wins assignment (:293-298) then scipy.ndimage.gaussian_filter
 *              (separable, mode="reflect", taps exp(-k^2 / 2 sigma^2) / sum for |k| <= radius; the caller
 *              passes radius = int(truncate * sigma + 0.5), scipy's rule with truncate = 4);
 *   fast == 0: _subpixel_gaussian (:314-359), additive, clip box, no border folding, for the spots inside the
 *              frame (what get_diffraction_pattern passes on, simulation2d.py:422-430);
 *   fast == 2: the same without the in-frame selection (the bare function: a spot centred outside the frame
 *              still spreads into it).
 * (DiffractionSimulation.get_diffraction_pattern, diffsims/sims/diffraction_simulation.py:296-354, is the
 * same computation for square shapes: pattern[x, y] = I followed by .T.)
 * images[n_tmpl][H][W] float32.  Rows of xyz are [cap][3] doubles (z ignored).
 */
int64_t ds_render_scratch_bytes(int32_t n_tmpl, int32_t cap);
/* number of kernels ds_render launches for a configuration (1, or 2 when the tcgen05 path runs its per-template
   prepare pass first); the dispatch rule itself, exported for launch accounting */
int ds_render_launch_count(int32_t cap, int32_t H, int32_t W, int32_t radius, int32_t fast, double mean_spots_hint);
int ds_render(void *stream,
              int32_t n_tmpl, int32_t cap, const int32_t *count /*[n_tmpl]*/,
              const double *xyz /*[n_tmpl][cap][3]*/, const double *intensity /*[n_tmpl][cap]*/,
              int32_t H, int32_t W,
              double calibration, double cx, double cy, double in_plane_angle_deg, int32_t mirrored,
              int32_t fast, double sigma, int32_t radius,
              double clip_threshold, int32_t normalize,
              float *images /*[n_tmpl][H][W]*/,
              void *scratch /*device scratch of ds_render_scratch_bytes(n_tmpl, cap) bytes, 16-byte aligned, contents
                              irrelevant on entry; one buffer per concurrently used stream.  Holds the ticket words
                              that hand templates to CTAs dynamically and, for the tcgen05 path, one prepared record
                              per template (live spots sorted by detector column + per-half lists)*/,
              double mean_spots_hint /*expected reflections per template, <= 0 if unknown: only tunes the
                                       schedule (front warps per CTA), never the result*/);

/*
 * Polar flattening of the packed result for template matching.
 * Replaces Simulation2D.polar_flatten_simulations (diffsims/simulations/simulation2d.py:313-355) with
 * DiffractingVector.to_flat_polar (diffsims/crystallography/_diffracting_vector.py:186-194) and get_closest
 * (simulation2d.py:767-781): per template r = |g_xy|, theta = atan2(y, x), intensity, zero padded to
 * max_spots.  With axes (both or neither) r and theta are replaced by the index of the closest axis entry
 * and spots with r_idx >= n_radial - 1 or theta_idx >= n_azimuthal - 1 are dropped (then compacted).
 * Outputs are float64 [n_tmpl][max_spots] (indices are stored as integral doubles).
 */
int ds_polar_flatten(void *stream, int32_t n_tmpl, int32_t cap, const int32_t *count /*[n_tmpl]*/,
                     const double *xyz /*[n_tmpl][cap][3]*/, const double *intensity /*[n_tmpl][cap]*/,
                     int32_t max_spots, int32_t n_radial, const double *radial_axes /*[n_radial] or NULL*/,
                     int32_t n_azimuthal, const double *azimuthal_axes /*[n_azimuthal] or NULL*/,
                     double *r_out, double *theta_out, double *intensity_out);

/*
 * Padded per-rotation rows (the output of ds_simulate) -> CSR: template t's min(count[t], cap) reflections are
 * copied to offsets[t] .. offsets[t + 1] - 1 of the packed arrays (offsets: int64 [n_tmpl + 1], the exclusive
 * prefix sums of the counts).  This is the "packed spot list" form of a library: what a sharded build gathers
 * at its end (the loops it replaces are diffsims/generators/simulation_generator.py:198,
 * :211 and diffsims/generators/library_generator.py:107, :117) and what a host consumer of
 * calculate_diffraction2d receives.  Any of the three outputs may be NULL.
 */
int ds_pack_csr(void *stream, int32_t n_tmpl, int32_t cap, const int32_t *count /*[n_tmpl]*/,
                const int64_t *offsets /*[n_tmpl + 1]*/, const int32_t *g_index /*[n_tmpl][cap]*/,
                const double *xyz /*[n_tmpl][cap][3]*/, const double *intensity /*[n_tmpl][cap]*/,
                int32_t *g_index_out /*[total]*/, double *xyz_out /*[total][3]*/, double *intensity_out /*[total]*/);

/*
 * Optional 16-bit export of NORMALISED templates (values in [0, 1], what Simulation2D.get_diffraction_pattern returns
 * with the reference's default normalisation, diffsims/simulations/simulation2d.py:440-441):
 * out[i] = rint(clamp(in[i], 0, 1) * 65535).  The quantisation step (1.5e-5 of the peak) is below the 1e-4 parity
 * budget of the float32 templates; a host consumer that accepts uint16 halves the device->host bytes.  Both buffers
 * 16-byte aligned.  float32 stays the product of ds_render; this is a separate, optional pass.
 */
int ds_quantize_u16(void *stream, int64_t n, const float *in /*[n]*/, uint16_t *out /*[n]*/);

/*
 * Pixel coordinates of an old-api template library: rint((xy + offset) / calibration + half_shape) as int32
 * (diffsims/generators/library_generator.py:129-132, diffsims/sims/diffraction_simulation.py:143-149).
 * pixel_coords[n_tmpl][cap][2]; entries beyond count[t] are zero.
 */
int ds_library_pixel_coords(void *stream, int32_t n_tmpl, int32_t cap, const int32_t *count /*[n_tmpl]*/,
                            const double *xyz /*[n_tmpl][cap][3]*/, double calibration_x, double calibration_y,
                            double offset_x, double offset_y, double half_shape_x, double half_shape_y,
                            int32_t *pixel_coords);

/*
 * Rotation-list producer: beam directions inside the stereographic triangle of a crystal system as Bunge
 * Euler angles (0, Phi, phi2) in degrees and/or as the active quaternions ds_simulate consumes.
 * Replaces get_beam_directions_grid (diffsims/generators/rotation_list_generators.py:176-267) for the cube
 * meshes of get_cube_mesh_vertices (diffsims/generators/sphere_mesh_generators.py:96-197) with
 * beam_directions_grid_to_euler (:486-526).  i_vals[n_i] is the 1-D face grid (device, computed by the host
 * as the reference does: np.tan spacing per mesh type); mode 0 = no crop (triclinic), 1 = x >= epsilon
 * (monoclinic as the reference evaluates it), 2 = the three plane tests n_k . v >= epsilon with
 * normals_host[3][3] (HOST pointer).  Order-preserving compaction in two passes over
 * ds_beam_grid_num_blocks(n_i) blocks: pass 0 fills block_counts, the caller exclusive-scans them into
 * block_offsets (int64), pass 1 writes the survivors.  euler_deg / quat_active may be NULL.
 */
int64_t ds_beam_grid_num_blocks(int32_t n_i);
int ds_beam_grid(void *stream, int32_t pass, int32_t n_i, const double *i_vals, int32_t mode,
                 const double *normals_host /*[3][3] or NULL*/, double epsilon, int32_t *block_counts,
                 const int64_t *block_offsets, double *euler_deg /*[n][3]*/, double *quat_active /*[n][4]*/);

/*
 * The same crop / ordered compaction / Euler + quaternion conversion for mesh vertices already in device memory
 * (points[n_points][3], float64): the uv-sphere, icosahedral and random meshes of
 * diffsims/generators/sphere_mesh_generators.py:42-93, :378-483 selected by get_beam_directions_grid(mesh=...)
 * (rotation_list_generators.py:205-235).  Passes and block arrays as for ds_beam_grid, over
 * ds_beam_points_num_blocks(n_points) blocks.
 */
int64_t ds_beam_points_num_blocks(int64_t n_points);
int ds_beam_points(void *stream, int32_t pass, int64_t n_points, const double *points, int32_t mode,
                   const double *normals_host, double epsilon, int32_t *block_counts,
                   const int64_t *block_offsets, double *euler_deg, double *quat_active);

/*
 * Rotation-list producer over SO(3): the cubochoric equal-volume grid of rotations ((2 n_steps)^3 cell-centred points
 * of the cube of edge pi^(2/3), mapped to unit quaternions) cropped to the fundamental zone of a proper point group
 * (mode 1: sym_quats_host[n_sym][4], HOST pointer, <= 24 operations; a rotation is kept iff none of its symmetric
 * equivalents has a smaller rotation angle), to rotation angles <= max_angle_rad (mode 2) or not at all (mode 0),
 * optionally composed with a centre rotation (centre_quat_host[4] or NULL: q_out = centre * q).
 * Replaces get_fundamental_zone_grid / get_local_grid (diffsims/generators/rotation_list_generators.py:85-134), i.e.
 * orix.sampling.get_sample_fundamental / get_sample_local (third party, source absent: parity with orix's point
 * lists is UNPINNED; the algorithm is the published one, see csrc/so3_grid.cu).  Order-preserving compaction in two
 * passes over ds_so3_grid_num_blocks(n_steps) blocks exactly as ds_beam_grid.  euler_deg [n][3] (Bunge, degrees) and
 * quat_active [n][4] (the conjugates, what ds_simulate consumes) may be NULL.
 */
int64_t ds_so3_grid_num_blocks(int32_t n_steps);
int ds_so3_grid(void *stream, int32_t pass, int32_t n_steps, int32_t mode, int32_t n_sym,
                const double *sym_quats_host, double max_angle_rad, const double *centre_quat_host,
                int32_t *block_counts, const int64_t *block_offsets, double *euler_deg, double *quat_active);

#ifdef __cplusplus
}
#endif
#endif /* DIFFSIMS_B200_H */
