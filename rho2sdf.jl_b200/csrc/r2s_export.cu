// r2s_export.cu -- exportSdfToVTI (DataExport/ExportToVTI.jl:22-67) for results that live on the device or on the host.
//
// at 10^9 voxels the writer dominates the wall time of rho2sdf() after the timed region.  The
// reference hands a 3-D array to WriteVTK (vtk_grid + vtk_point_data + vtk_save).  Here the same VTK ImageData file is
// written directly: an XML header (WholeExtent = 0..dims-1, Origin = grid.AABB_min, Spacing = cell_size / smooth, one
// PointData scalar named by the caller, "distance" in rho2sdf) followed by ONE raw appended block, little endian, UInt64
// length header -- readable by ParaView / VTK without decompression.  Device-resident results are streamed plane-chunk by
// plane-chunk (device -> pinned staging buffer -> file), so no host copy of the whole field is ever needed.
#include <stdio.h>
#include <string.h>
#include <string>
#include <algorithm>
#include "r2s_common.cuh"

// XML attribute values: the caller's label is escaped (a quote, '<' or '&' in it must not break the header)
static std::string xml_escape(const char *t) {
  std::string o;
  for (const char *p = t; *p; p++) {
    switch (*p) { case '&': o += "&amp;"; break; case '<': o += "&lt;"; break; case '>': o += "&gt;"; break; case '"': o += "&quot;"; break; case '\'': o += "&apos;"; break; default: o += *p; }
  }
  return o;
}
// z0..z1: the planes this file holds (a piece of a slab decomposition holds a sub-range; a whole-grid file 0..nz-1)
static int write_header(FILE *f, const char *label, int is_f64, i64 nx, i64 ny, i64 z0, i64 z1, const double origin[3], const double spacing[3]) {
  const std::string lab = xml_escape(label);
  int n = fprintf(f,
                  "<?xml version=\"1.0\"?>\n<VTKFile type=\"ImageData\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n"
                  "  <ImageData WholeExtent=\"0 %lld 0 %lld %lld %lld\" Origin=\"%.17g %.17g %.17g\" Spacing=\"%.17g %.17g %.17g\">\n"
                  "    <Piece Extent=\"0 %lld 0 %lld %lld %lld\">\n      <PointData Scalars=\"%s\">\n"
                  "        <DataArray type=\"%s\" Name=\"%s\" NumberOfComponents=\"1\" format=\"appended\" offset=\"0\"/>\n"
                  "      </PointData>\n      <CellData/>\n    </Piece>\n  </ImageData>\n  <AppendedData encoding=\"raw\">\n_",
                  nx - 1, ny - 1, z0, z1, origin[0], origin[1], origin[2], spacing[0], spacing[1], spacing[2], nx - 1, ny - 1, z0, z1, lab.c_str(),
                  is_f64 ? "Float64" : "Float32", lab.c_str());
  if (n < 0) return 1;
  uint64_t bytes = (uint64_t)nx * ny * (uint64_t)(z1 - z0 + 1) * (is_f64 ? 8 : 4);
  return fwrite(&bytes, sizeof(bytes), 1, f) == 1 ? 0 : 1;
}
static int write_footer(FILE *f) { return fputs("\n  </AppendedData>\n</VTKFile>\n", f) < 0 ? 1 : 0; }

extern "C" {

// host data: values[nx*ny*nz], x fastest (the layout of every field of this library and of Julia's reshape(values, dims...))
int r2s_write_vti_host(const char *path, const char *label, const void *values, int is_f64, int64_t nx, int64_t ny, int64_t nz, const double origin[3],
                       const double spacing[3]) {
  if (!path || !label || !values || nx < 1 || ny < 1 || nz < 1) return 1;
  FILE *f = fopen(path, "wb");
  if (!f) return 2;
  int rc = write_header(f, label, is_f64, nx, ny, 0, nz - 1, origin, spacing);
  size_t n = (size_t)nx * ny * nz, es = is_f64 ? 8 : 4;
  if (!rc && fwrite(values, es, n, f) != n) rc = 3;
  if (!rc) rc = write_footer(f);
  if (fclose(f) != 0 && !rc) rc = 4;
  return rc;
}

// device-resident result of the last pipeline / smoothing call: which = 0 -> sdf_dists (Float64, coarse grid), 1 -> fine_sdf
// (Float32, dims N*smooth+1).  On a slab rank (multi-GPU) the call is COLLECTIVE and writes this rank's PIECE: its own planes plus
// the first plane of the upper neighbour (VTK pieces share their boundary points), fetched with one halo exchange;
// r2s_export_pvti writes the index file that ties the pieces together.
int r2s_export_vti(r2s_ctx *ctx, const char *path, const char *label, int which) {
  if (!ctx) return 1;
  if (!path || !label) FAIL("r2s_export_vti: path and label are required");
  if (!ctx->has_grid) FAIL("r2s_set_grid has not been called");
  if (which == 1 ? !ctx->have_fine : !ctx->have_sdf) FAIL("r2s_export_vti: no result for the current grid on the device yet (run the pipeline first)");
  const GridDev &g = ctx->g;
  const int s = which == 1 ? ctx->smooth_last : 1;
  const i64 nx = g.N[0] * (i64)s + 1, ny = g.N[1] * (i64)s + 1, nz = g.N[2] * (i64)s + 1;
  const size_t es = which == 1 ? 4 : 8, plane = (size_t)nx * ny * es;
  const char *src = which == 1 ? (const char *)ctx->f_fine.p : (const char *)ctx->sdf.p;
  CK(cudaSetDevice(ctx->device));
  // planes of this file
  i64 z0 = 0, z1 = nz - 1;
  if (ctx->nranks > 1) {
    const i64 k0 = ctx->k0, k1 = ctx->k1, nzc = g.np[2];
    z0 = s * k0; z1 = (k1 < nzc) ? s * k1 : nz - 1;      // inclusive; the plane s*k1 belongs to the upper neighbour
    const i64 own1 = (k1 < nzc) ? s * k1 : nz;
    if (r2s_halo_exchange_f32(ctx, (float *)src, (i64)(plane / 4), (int)z0, (int)own1, (int)nz, 0, 1)) return 1;
  }
  const double origin[3] = {g.amin[0], g.amin[1], g.amin[2]}, spacing[3] = {g.cell / s, g.cell / s, g.cell / s};      // ExportToVTI.jl:31-42
  FILE *f = fopen(path, "wb");
  if (!f) FAIL(std::string("r2s_export_vti: cannot open ") + path);
  int rc = write_header(f, label, which != 1, nx, ny, z0, z1, origin, spacing);
  // two pinned staging buffers of ~64 MB: the copy of chunk c+1 overlaps the fwrite of chunk c
  const i64 cpl = std::max<i64>(1, (i64)((64u << 20) / plane));
  void *stage[2] = {nullptr, nullptr};
  if (!rc && (cudaMallocHost(&stage[0], plane * cpl) != cudaSuccess || cudaMallocHost(&stage[1], plane * cpl) != cudaSuccess)) rc = 5;
  if (!rc) {
    i64 k = z0; int b = 0;
    i64 n0 = std::min<i64>(cpl, z1 + 1 - z0);
    if (cudaMemcpyAsync(stage[0], src + plane * (size_t)z0, plane * n0, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = 6;
    while (!rc && k <= z1) {
      const i64 nk = std::min<i64>(cpl, z1 + 1 - k);
      if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = 6; break; }
      const i64 k2 = k + nk;
      if (k2 <= z1) {
        const i64 n2 = std::min<i64>(cpl, z1 + 1 - k2);
        if (cudaMemcpyAsync(stage[1 - b], src + plane * (size_t)k2, plane * n2, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) { rc = 6; break; }
      }
      if (fwrite(stage[b], 1, plane * nk, f) != plane * nk) { rc = 3; break; }
      k = k2; b = 1 - b;
    }
  }
  cudaStreamSynchronize(ctx->stream);
  if (stage[0]) cudaFreeHost(stage[0]);
  if (stage[1]) cudaFreeHost(stage[1]);
  if (!rc) rc = write_footer(f);
  if (fclose(f) != 0 && !rc) rc = 4;
  if (rc) { char b[128]; snprintf(b, sizeof(b), "r2s_export_vti failed (code %d)", rc); FAIL(b); }
  return 0;
}
// Index file of a slab decomposition (VTK PImageData): one <Piece> per slab, Source = "<piece_base>_<slab>.vti" (file names relative to
// the directory of the .pvti).  Any rank may call it (every rank knows the partition); it touches no device memory.
int r2s_export_pvti(r2s_ctx *ctx, const char *path, const char *label, int which, const char *piece_base) {
  if (!ctx) return 1;
  if (!path || !label || !piece_base) FAIL("r2s_export_pvti: path, label and piece_base are required");
  if (!ctx->has_grid || ctx->nranks < 2 || (int)ctx->slab_k0.size() != ctx->nranks + 1) FAIL("r2s_export_pvti: no slab decomposition on this context");
  const GridDev &g = ctx->g;
  const int s = which == 1 ? ctx->smooth_last : 1;
  const i64 nx = g.N[0] * (i64)s + 1, ny = g.N[1] * (i64)s + 1, nz = g.N[2] * (i64)s + 1;
  std::string pb(piece_base); const size_t sl = pb.find_last_of('/'); if (sl != std::string::npos) pb = pb.substr(sl + 1);
  if (pb.size() > 4 && pb.substr(pb.size() - 4) == ".vti") pb.resize(pb.size() - 4);
  const std::string lab = xml_escape(label), src = xml_escape(pb.c_str());
  FILE *f = fopen(path, "wb");
  if (!f) FAIL(std::string("r2s_export_pvti: cannot open ") + path);
  fprintf(f, "<?xml version=\"1.0\"?>\n<VTKFile type=\"PImageData\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n"
             "  <PImageData WholeExtent=\"0 %lld 0 %lld 0 %lld\" GhostLevel=\"0\" Origin=\"%.17g %.17g %.17g\" Spacing=\"%.17g %.17g %.17g\">\n"
             "    <PPointData Scalars=\"%s\">\n      <PDataArray type=\"%s\" Name=\"%s\" NumberOfComponents=\"1\"/>\n    </PPointData>\n",
          nx - 1, ny - 1, nz - 1, g.amin[0], g.amin[1], g.amin[2], g.cell / s, g.cell / s, g.cell / s, lab.c_str(), which != 1 ? "Float64" : "Float32", lab.c_str());
  for (int r = 0; r < ctx->nranks; r++) {
    const i64 k0 = ctx->slab_k0[(size_t)r], k1 = ctx->slab_k0[(size_t)r + 1];
    fprintf(f, "    <Piece Extent=\"0 %lld 0 %lld %lld %lld\" Source=\"%s_%d.vti\"/>\n", nx - 1, ny - 1, s * k0, (k1 < g.np[2]) ? s * k1 : nz - 1, src.c_str(), r);
  }
  const int bad = fputs("  </PImageData>\n</VTKFile>\n", f) < 0;
  if (fclose(f) != 0 || bad) FAIL("r2s_export_pvti: write failed");
  return 0;
}
}  // extern "C"
