/* oracle/linkproof/hdf5.h -- TEST INFRASTRUCTURE.  libhdf5 is not in this image; the link proof (oracle/Makefile,
 * target `linkproof`) compiles the reference's src/flappie.c UNMODIFIED, and that file only needs these names from
 * <hdf5.h> (src/flappie.c:325,380,390-392; src/fast5_interface.h:19-23). */
#ifndef FFB_LINKPROOF_HDF5_H
#define FFB_LINKPROOF_HDF5_H
typedef long hid_t;
typedef unsigned long long hsize_t;
static inline int H5Fclose(hid_t h) { (void)h; return 0; }
#endif
