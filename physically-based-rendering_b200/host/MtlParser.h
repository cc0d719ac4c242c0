/*
 * MtlParser -- reference: source/MtlParser.{h,cpp}.  Same struct material_t, same class surface
 * (getMaterials, load), same keys incl. the custom ones (light, rough, p, nu, nv, Rs, Rd) and the
 * same quirks: lines shorter than 3 characters are skipped, `Tr` is ignored once ANY `d` has been
 * seen in the file (MtlParser.cpp:57, 99-108), values are read with atof.
 */
#ifndef MTLPARSER_H
#define MTLPARSER_H

#include <string>
#include <vector>

#include "cl_types.h"
#include "Logger.h"

struct material_t {
	std::string mtlName;
	cl_float4 Ka, Kd, Ks;        /* ambient (unused by the kernel), diffuse, specular colour */
	cl_float d, Ni, Ns;          /* opacity, index of refraction, Phong exponent (unused) */
	cl_char illum;
	cl_char light;               /* 1 = emitter; parsed, ignored by the kernel */
	cl_float rough, p;           /* Schlick: roughness, isotropy */
	cl_float nu, nv, Rs, Rd;     /* Shirley-Ashikhmin: lobe exponents, specular / diffuse weight */
};

class MtlParser {
	public:
		void load( std::string file );
		std::vector<material_t> getMaterials();
		static material_t getEmptyMaterial();           /* the defaults of MtlParser.cpp:11-36 */
		/** Additive: install materials that were not read from a file (synthetic scenes). */
		void setMaterials( const std::vector<material_t>& materials );

	private:
		std::vector<material_t> mMaterials;
};

#endif
