// tb_shape.cu -- one board shape's kernels and launchers: compiled once per shape with -DTB_C=<columns> -DTB_R=<rows>.
// Exports `const TbShapeVT *tb_shape_vt_<C>x<R>(void)` (and the same table as `tb_shape_vt` when built as a stand-alone
// shape object with -DTB_SHAPE_PLUGIN, for tb_load_shape).
#include "tb_kernels.cuh"

#if !defined(TB_C) || !defined(TB_R)
#error "compile with -DTB_C=<num_columns> -DTB_R=<num_rows>"
#endif

#define TB_CAT_(a, b, c, d) a##b##c##d
#define TB_CAT(a, b, c, d) TB_CAT_(a, b, c, d)

extern "C" const TbShapeVT *TB_CAT(tb_shape_vt_, TB_C, x, TB_R)(void) { return tb::ShapeOps<TB_C, TB_R>::vt(); }
#ifdef TB_SHAPE_PLUGIN
extern "C" const TbShapeVT *tb_shape_vt(void) { return tb::ShapeOps<TB_C, TB_R>::vt(); }
#endif
