/* empty stand-in: <io.h> is a Windows header the reference's threadpool.h includes */
